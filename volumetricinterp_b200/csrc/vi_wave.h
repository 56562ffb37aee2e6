// Rotation-tape replay by ONE WARP per system, sweeps pipelined across the lanes (kernel K3c).
//
// The QL kernel (vi_tql.h) leaves, per eigen-system, a tape of ~8 000 Givens rotations (n = 144); the truncated
// spectral solve needs  Z^T g  (tape forwards)  and  Z u  (tape backwards).  A thread-per-system replay is one long
// dependent chain -- every rotation shares an element with the next -- and ran at ~320 cycles per rotation and lane.
// But the tape is a sequence of QL SWEEPS, each a monotone walk over adjacent planes (i, i+1), (i-1, i), ...: sweep
// t+1 can rotate plane p as soon as sweep t is done with the planes p-1, p, p+1.  So lane L of the warp takes the
// sweeps L, L+32, ... and follows its predecessor (the sweep on lane L-1) two planes behind: a diagonal wavefront of
// up to 32 rotations per step on disjoint pairs of elements of the vector, which lives in shared memory.
//
// Sweep boundaries are recovered from the tape itself (inside a sweep the plane index moves by exactly one per
// rotation, in the direction the entry's flag gives; a new sweep never starts one plane further on), so the QL kernel
// and the tape format are unchanged.  Backwards (Z u) the same wavefront runs over the sweeps in reverse order, each
// walked from its last rotation to its first.
//
// Reference call being replaced: scipy.linalg.lstsq at interpolate.py:462 (see vi_tql.h).
#pragma once
#include "vi_simt.h"

#ifndef VI_HD
#if defined(__CUDACC__)
#define VI_HD __host__ __device__ __forceinline__
#else
#define VI_HD inline
#endif
#endif

// sweep table: start offsets in the tape, capacity per system (QL makes ~1.7 sweeps per eigenvalue; its iteration
// budget is 30 n -- a tape with more sweeps than this is replayed in several table loads)
VI_HD int vi_wav_maxsweeps(int n) { return 4 * n + 32; }
// shared-memory bytes per warp: vector (n doubles, padded) + sweep table; without the table when it lives in global memory
VI_HD int vi_wav_bytes(int n) { return (((n + 7) & ~7) * 8 + 64 + (vi_wav_maxsweeps(n) + 1) * 4 + 15) & ~15; }   // warps stay 16-byte aligned
VI_HD int vi_wav_bytes_vec(int n) { return ((n + 7) & ~7) * 8 + 64; }

#if defined(__CUDACC__) || defined(VI_EMU)

VI_DEV int vi_ballot(bool p) {
#if defined(__CUDACC__)
  return (int)__ballot_sync(0xffffffffu, p);
#else
  // emulator: one bit per lane through the integer mailbox
  int m = 0;
  for (int l = 0; l < 32; ++l) m |= vi_shfl_i(p ? 1 : 0, l) << l;
  return m;
#endif
}
VI_DEV int vi_popc(unsigned x) {
#if defined(__CUDACC__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// one tape entry's (c, s): a single 16-byte read-only load on the device
VI_DEV void vi_wav_ld(const double* cs, int t, double& c, double& s) {
#if defined(__CUDACC__)
  const double2 v = __ldg(reinterpret_cast<const double2*>(cs) + t);
  c = v.x; s = v.y;
#else
  c = cs[2 * t]; s = cs[2 * t + 1];
#endif
}

// Sweep starts of tape entries [t0, nrot): an entry continues the sweep of its predecessor iff it has the same
// direction flag and its plane index is exactly one step further on.  Fills tab[0 .. ns] (tab[ns] = end offset of the
// last complete sweep taken, or nrot), at most cap sweeps; returns ns.  *tnext = where the next table load starts.
VI_DEV int vi_wav_scan(const int32_t* ix, int t0, int nrot, int32_t* tab, int cap, int* tnext) {
  const int lane = vi_tid() & 31;
  int ns = 0;
  int prev = 0;                                  // entry t - 1 (carried across the 32-entry trips)
  int stop = nrot;
  for (int base = t0; base < nrot; base += 32) {
    const int t = base + lane;
    const int code = (t < nrot) ? ix[t] : 0;
    int before = vi_shfl_i(code, (lane + 31) & 31);
    if (lane == 0) before = prev;
    const int step = (code & 1) ? 2 : -2;        // plane index (code >> 1) moves by +1 (reversed block) or -1 per rotation
    const bool start = (t < nrot) && (t == t0 || ((code ^ before) & 1) != 0 || code != before + step);
    const unsigned m = (unsigned)vi_ballot(start);
    const int pos = ns + vi_popc(m & ((1u << lane) - 1u));
    if (start && pos <= cap) tab[pos] = t;       // entry `cap` (if reached) is the end marker of a full table
    ns += vi_popc(m);
    prev = vi_shfl_i(code, 31);
    if (ns > cap) { stop = -1; break; }
  }
  vi_warp_sync();
  if (stop < 0) {                                // table full: sweep `cap` starts at tab[cap]; it is not taken
    ns = cap;
    *tnext = tab[cap];
  } else {
    if (lane == 0) tab[ns] = nrot;
    *tnext = nrot;
  }
  vi_warp_sync();
  return ns;
}

// One wavefront pass over the sweeps tab[0 .. ns).  FWD: w <- Z^T w (sweeps in order, rotations in order, as
// vi_tape_apply_zt); otherwise w <- Z w (sweeps in reverse order, rotations of each in reverse, as vi_tape_apply_z).
// w: n doubles in shared memory.  cs: (c, s) pairs, ix: codes (vi_tql.h).
//
// Ordering rule.  Sweep q (in processing order) walks its planes in direction u = +-1; a rotation at plane p touches
// the elements p and p -+ 1 behind it.  Let B_q be the FRONTIER of sweep q: every element on the far side of B_q (in
// direction -u) is final with respect to ALL sweeps <= q.  B_q = min over the chain of (own position - 2 u), carried
// in the key K = u B so that both directions read "rotate plane p iff u p <= K_{q-1}".  Each lane publishes
// (q, u, K, all-earlier-sweeps-done) and reads its predecessor's -- the previous lane's -- values of the step before
// (stale values are smaller, hence safe).  A lane opens its next sweep only once every earlier sweep is done, so a
// predecessor lane that has moved on means "everything before me is finished".  Sweeps whose ranges do not nest
// (short sweeps on a split-off sub-block followed by a long one) are ordered by the same rule: the frontier is
// cumulative along the chain, not the predecessor's position alone.
//
// G = lanes per system (32, 16 or 8): a warp replays 32 / G systems side by side, each lane group with its own vector,
// tape and sweep table (the arguments are per lane, uniform inside a group).  The sweeps of a QL run are short on
// average (7 of 32 lanes rotate per step with G = 32): at full batches four systems per warp cost a warp little more
// than one.
template <bool FWD, int G = 32>
VI_DEV void vi_wav_pass(double* w, const double* cs, const int32_t* ix, const int32_t* tab, int ns) {
  const int lane = vi_tid() & 31;
  const int gl = lane & (G - 1), gbase = lane & ~(G - 1);
  const int BIG = 1 << 28;
  const int dt = FWD ? 1 : -1;
  int q = gl;                                    // position in processing order: sweep FWD ? q : ns - 1 - q
  int k = 0, len = 0, t = 0;                     // progress inside the sweep, its length, current tape entry
  int pi = 0, u = 1, flag = 0;                   // plane index of the current rotation, its step per rotation, direction flag
  double c = 1.0, s = 0.0, c1 = 1.0, s1 = 0.0;   // (c, s) of the current entry and of the one after it (load in flight;
                                                 // three more in flight measured slower: 9.0 vs 8.7 ms per 28 416 systems)
  // Everything a lane needs to OPEN a sweep is fetched one or two sweeps ahead (a sweep open is otherwise two dependent
  // global loads -- table, then the first tape entry -- in front of every lane of the warp): p1 = next sweep of this
  // lane (table entries + first tape entry), p2 = the one after (table entries).  Inside a sweep only (c, s) are read:
  // the plane index moves by u per rotation and the direction flag is constant (that is what defines a sweep).
  int p1a = 0, p1b = 0, p1code = 0, p2a = 0, p2b = 0;
  double p1c = 1.0, p1s = 0.0;
  auto tabld = [&](int pos, int& a, int& b) {
    if (pos < ns) { const int sw = FWD ? pos : ns - 1 - pos; a = tab[sw]; b = tab[sw + 1]; }
  };
  auto firstld = [&](int pos, int a, int b, int& cd, double& cc, double& ss) {
    if (pos < ns) { const int tt = FWD ? a : b - 1; cd = ix[tt]; vi_wav_ld(cs, tt, cc, ss); }
  };
  auto start = [&]() {                           // sweep q from the p1 registers
    k = 0; len = 0;
    if (q < ns) {
      len = p1b - p1a;
      t = FWD ? p1a : p1b - 1;
      c = p1c; s = p1s;
      pi = p1code >> 1;
      flag = p1code & 1;
      const int along = flag ? 1 : -1;           // plane index step per rotation in tape order
      u = FWD ? along : -along;
      if (len > 1) vi_wav_ld(cs, t + dt, c1, s1);
    }
  };
  tabld(q, p1a, p1b);
  firstld(q, p1a, p1b, p1code, p1c, p1s);
  start();
  tabld(q + G, p1a, p1b);
  firstld(q + G, p1a, p1b, p1code, p1c, p1s);
  tabld(q + 2 * G, p2a, p2b);
  int K = -BIG;                                  // published frontier key of my current sweep (cumulative)
  int cum = 0;                                   // published: my sweep and every earlier one are done
  for (;;) {
    const bool have = q < ns;
    const bool done = !have || k >= len;
    // predecessor = processing position q - 1, on the previous lane of the group; its values as published last step
    const int pl = gbase + ((gl + G - 1) & (G - 1));
    const int pq = vi_shfl_i(q, pl), pu = vi_shfl_i(u, pl), pK = vi_shfl_i(K, pl), pcum = vi_shfl_i(cum, pl);
    int lim;                                     // K_{q-1} as far as I may rely on it
    bool before_done;                            // every sweep < q is finished
    if (q == 0) { lim = BIG; before_done = true; }
    else if (pq > q - 1) { lim = BIG; before_done = true; }                  // it moved on: all earlier ones are done
    else if (pq < q - 1) { lim = -BIG; before_done = false; }                // not opened yet
    else {
      before_done = pcum != 0;
      lim = before_done ? BIG : ((pu == u) ? pK : -BIG);
    }
    if (have && !done && u * pi <= lim) {
      const int pj = flag ? pi - 1 : pi + 1;
      const double a = w[pi], b = w[pj];
      if (FWD) { w[pj] = s * a + c * b; w[pi] = c * a - s * b; }
      else { w[pi] = c * a + s * b; w[pj] = c * b - s * a; }
      ++k; t += dt; pi += u;
      c = c1; s = s1;
      if (k + 1 < len) vi_wav_ld(cs, t + dt, c1, s1);
    }
    const bool fin = !have || k >= len;
    // publish; a finished sweep hands the chain's frontier on unchanged
    {
      const int own = fin ? BIG : u * pi - 2;
      K = (lim < own) ? lim : own;
      cum = (fin && before_done) ? 1 : 0;
    }
    if (have && fin && before_done) {            // next sweep of this lane
      q += G;
      start();
      p1a = p2a; p1b = p2b;
      firstld(q + G, p1a, p1b, p1code, p1c, p1s);
      tabld(q + 2 * G, p2a, p2b);
      K = -BIG; cum = 0;                         // nothing known yet about the new sweep's chain
    }
    vi_warp_sync();
    if (vi_ballot(q < ns) == 0) break;           // every lane has run out of sweeps
  }
}

#endif  // device / emulator
