// TEST-ONLY: a small CUDA-on-CPU execution model, just enough to run the warp-synchronous device code of
// volumetricinterp_b200/csrc (vi_band.h, vi_chase.h) UNCHANGED in the GPU-less build container.
//
// One CTA at a time.  Every CUDA thread is a fiber (ucontext) on one OS thread, scheduled round-robin; a fiber
// runs until it reaches a synchronisation point (__syncthreads, __syncwarp, a shuffle, an mma.sync) and yields
// until all participants have arrived.  Shuffles and the m8n8k4 DMMA exchange their operands through a per-warp
// mailbox.  Deterministic (no races), slow (microseconds per synchronisation point), never part of the product.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <functional>
#include <vector>

namespace emu {

struct Dim3 { unsigned x = 1, y = 1, z = 1; };

struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  int tid = 0;
  bool done = false;
};

struct Barrier { int arrived = 0; uint64_t gen = 0; };

struct Warp {
  Barrier bar;
  double box[32];
  double boxa[32], boxb[32];
};

inline uint64_t& progress() { static uint64_t p = 0; return p; }    // bumped whenever any barrier completes

struct Cta {
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  Barrier named[16];
  ucontext_t sched;
  int cur = 0;
  int nthreads = 0;
  unsigned block = 0;
  std::function<void()> body;
};

inline Cta*& cta() { static Cta* c = nullptr; return c; }
inline int tid() { return cta()->fibers[cta()->cur].tid; }

inline void yield() {
  Cta* c = cta();
  swapcontext(&c->fibers[c->cur].ctx, &c->sched);
}

inline void wait(Barrier& b, int count) {
  const uint64_t g = b.gen;
  if (++b.arrived == count) { b.arrived = 0; ++b.gen; ++progress(); return; }
  while (b.gen == g) yield();
}

inline void trampoline() {
  Cta* c = cta();
  c->body();
  c->fibers[c->cur].done = true;
  swapcontext(&c->fibers[c->cur].ctx, &c->sched);
}

// run one CTA of `nthreads` threads; body() is the kernel body (reads threadIdx through emu::tid())
inline void run_cta(unsigned block, int nthreads, const std::function<void()>& body) {
  Cta c;
  cta() = &c;
  c.nthreads = nthreads;
  c.block = block;
  c.body = body;
  c.fibers.resize(nthreads);
  c.warps.resize((nthreads + 31) / 32);
  for (int t = 0; t < nthreads; ++t) {
    Fiber& f = c.fibers[t];
    f.tid = t;
    f.stack.resize(256 * 1024);
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack.data();
    f.ctx.uc_stack.ss_size = f.stack.size();
    f.ctx.uc_link = &c.sched;
    makecontext(&f.ctx, (void (*)())trampoline, 0);
  }
  int live = nthreads;
  int idle_passes = 0;
  while (live > 0) {
    const uint64_t before = progress();
    const int live_before = live;
    for (int t = 0; t < nthreads; ++t) {
      if (c.fibers[t].done) continue;
      c.cur = t;
      swapcontext(&c.sched, &c.fibers[t].ctx);
      if (c.fibers[t].done) { --live; }
    }
    // a pass over all fibers in which no barrier completed and no fiber finished: nobody can ever move again
    // (e.g. a warp-wide shuffle reached by only part of the warp -- a divergent __shfl_sync on the GPU)
    if (progress() == before && live == live_before) {
      if (++idle_passes > 4) { fprintf(stderr, "cuda_emu: DEADLOCK (divergent synchronisation?)\n"); abort(); }
    } else {
      idle_passes = 0;
    }
  }
  cta() = nullptr;
}

// ---- the CUDA surface the device headers use ----------------------------------------------------------------
struct ThreadIdx { operator int() const { return tid(); } };
struct Idx3 { int x_() const { return tid(); } };

inline int lane() { return tid() & 31; }
inline Warp& warp() { return cta()->warps[tid() >> 5]; }
inline int warp_count(int w) {          // threads of warp w (the last warp may be partial)
  const int n = cta()->nthreads - 32 * w;
  return n > 32 ? 32 : n;
}

inline void syncthreads() { wait(cta()->named[0], cta()->nthreads); }
inline void named_barrier(int id, int count) { wait(cta()->named[id], count); }
inline void syncwarp() { wait(warp().bar, warp_count(tid() >> 5)); }

inline double shfl(double v, int src) {
  Warp& w = warp();
  const int n = warp_count(tid() >> 5);
  w.box[lane()] = v;
  wait(w.bar, n);
  const double r = w.box[src & 31];
  wait(w.bar, n);
  return r;
}
inline int shfl_i(int v, int src) { return (int)shfl((double)v, src); }

// D = A (8x4, row) * B (4x8, col) + C, fp64, fragment layout of mma.sync.aligned.m8n8k4.row.col.f64:
//   a: A[lane/4][lane%4]   b: B[lane%4][lane/4]   c/d: C[lane/4][2*(lane%4) + {0,1}]
inline void mma884(double& d0, double& d1, double a, double b, double c0, double c1) {
  Warp& w = warp();
  const int n = warp_count(tid() >> 5);
  const int l = lane();
  w.boxa[l] = a;
  w.boxb[l] = b;
  wait(w.bar, n);
  const int row = l >> 2, col = 2 * (l & 3);
  double s0 = c0, s1 = c1;
  for (int k = 0; k < 4; ++k) {
    const double av = w.boxa[row * 4 + k];
    s0 = fma(av, w.boxb[col * 4 + k], s0);
    s1 = fma(av, w.boxb[(col + 1) * 4 + k], s1);
  }
  wait(w.bar, n);
  d0 = s0; d1 = s1;
}

}  // namespace emu
