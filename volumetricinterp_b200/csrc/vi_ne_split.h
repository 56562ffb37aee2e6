// Work split of the version-3 normal-equation kernel (k_ne_dmma3, csrc/normal_eq.cu): host-side, plain C++, so that
// the GPU-less test harness can check it (tests/cpu_harness.cpp).
#pragma once

constexpr int kUnitDiag = 1 << 16;      // unit flag: diagonal unit (half tile + rhs)
constexpr int kUnitHalf = 1 << 17;      // ... whose half tile exists (8 (2 mi + 1) < N)
constexpr int kMaxUnits = 96;
constexpr int kW3 = 16;          // warps per CTA in version 3: four per tensor sub-partition, so that the ~4-deep
                                 // dependent DMMA chain of one accumulator is covered by the other warps' work
struct NeSplit {
  int nunits;
  int wbeg[kW3], wend[kW3];
  int uinfo[kMaxUnits];
};

// units in (mi, ni) order and their contiguous split over the warps by weight (full tile 4, diagonal unit 3):
// warp w ends where the running weight is closest to (w + 1) / kW3 of the total.  false if it does not fit.
inline bool ne3_split(int N, int mt, int DT, NeSplit& sp) {
  int c = 0, total = 0;
  for (int mi = 0; mi < mt; ++mi) {
    for (int ni = 0; ni <= 2 * mi; ++ni) {
      if (c >= kMaxUnits) return false;
      sp.uinfo[c++] = mi | (ni << 8); total += 4;
    }
    if (c >= kMaxUnits) return false;
    const bool half = 8 * (2 * mi + 1) < N;
    sp.uinfo[c++] = mi | ((2 * mi + 1) << 8) | kUnitDiag | (half ? kUnitHalf : 0);
    total += 3;
  }
  sp.nunits = c;
  // contiguous chunks of nearly equal weight ...
  int cbeg[kW3 + 1], cw[kW3];
  int u = 0, run = 0;
  for (int w = 0; w < kW3; ++w) {
    cbeg[w] = u;
    const int before = run;
    const int goal = (total * (w + 1) + kW3 / 2) / kW3;
    while (u < c && (u - cbeg[w]) < DT) {
      const int wt = (sp.uinfo[u] & kUnitDiag) ? 3 : 4;
      if (w < kW3 - 1 && run + wt > goal && run + wt - goal > goal - run) break;
      run += wt; ++u;
    }
    cw[w] = run - before;
  }
  cbeg[kW3] = u;
  if (u != c) return false;
  // ... dealt to the warps so that the four tensor sub-partitions (warp % 4) carry equal sums: heaviest chunk
  // first, each to the lightest sub-partition that still has a free warp
  bool used[kW3];
  int spsum[4] = {0, 0, 0, 0}, spcnt[4] = {0, 0, 0, 0};
  for (int w = 0; w < kW3; ++w) used[w] = false;
  for (int n = 0; n < kW3; ++n) {
    int best = -1;
    for (int w = 0; w < kW3; ++w)
      if (!used[w] && (best < 0 || cw[w] > cw[best])) best = w;
    used[best] = true;
    int q = -1;
    for (int p = 0; p < 4; ++p)
      if (spcnt[p] < kW3 / 4 && (q < 0 || spsum[p] < spsum[q])) q = p;
    const int warp = q + 4 * spcnt[q];
    spsum[q] += cw[best]; ++spcnt[q];
    sp.wbeg[warp] = cbeg[best];
    sp.wend[warp] = cbeg[best + 1];
  }
  return true;
}

