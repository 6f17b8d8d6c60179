// Host-only check of the partition geometry behind the whole-CU launches (hmb200_one.cuh: one_cu_part / one_cu_pus) against
// the reference's table, TComDataCU::getPartIndexAndSize (TLibCommon/TComDataCU.cpp:1893-1931), restated here on its own:
// 2Nx2N, 2NxN, Nx2N, 2NxnU, 2NxnD, nLx2N, nRx2N (AMP only above 8x8), and for a 16x16 CU the 2Nx2N / 2NxN / Nx2N PUs of its
// four 8x8 children in z-order.  Also checks what the kernels rely on: every PU is a union of cells of the CU's 4 x 4
// (8x8 CU: 2 x 2) grid, and PU offsets are multiples of 4 (row parity relative to the PU = relative to the CU).
// Built and run by tests/test_host_logic.py with nvcc; no GPU needed (only host code runs).
#include <cstdio>
#include <vector>
#include "hmb200_one.cuh"

struct Pu { int x, y, w, h; };

static std::vector<Pu> reference_parts(int S, bool amp) {
  const int h2 = S / 2, q = S / 4;
  std::vector<Pu> v = {{0, 0, S, S}, {0, 0, S, h2}, {0, h2, S, h2}, {0, 0, h2, S}, {h2, 0, h2, S}};
  if (amp) {
    v.push_back({0, 0, S, q});      v.push_back({0, q, S, S - q});          // SIZE_2NxnU
    v.push_back({0, 0, S, S - q});  v.push_back({0, S - q, S, q});          // SIZE_2NxnD
    v.push_back({0, 0, q, S});      v.push_back({q, 0, S - q, S});          // SIZE_nLx2N
    v.push_back({0, 0, S - q, S});  v.push_back({S - q, 0, q, S});          // SIZE_nRx2N
  }
  return v;
}

int main() {
  int bad = 0;
  for (int S : {8, 16, 32, 64}) {
    std::vector<Pu> want = reference_parts(S, S > 8);
    if (S == 16)
      for (int c = 0; c < 4; c++)
        for (const Pu& p : reference_parts(8, false)) want.push_back({p.x + (c & 1) * 8, p.y + (c >> 1) * 8, p.w, p.h});
    if ((int)want.size() != hmb200::one_cu_pus(S)) { printf("S %d: %zu PUs expected, one_cu_pus says %d\n", S, want.size(), hmb200::one_cu_pus(S)); bad++; continue; }
    const int nb = S == 8 ? 2 : 4, bs = S / nb;
    for (int p = 0; p < (int)want.size(); p++) {
      int x, y, w, h;
      hmb200::one_cu_part(S, p, &x, &y, &w, &h);
      if (x != want[p].x || y != want[p].y || w != want[p].w || h != want[p].h) {
        printf("S %d part %d: (%d,%d,%dx%d), reference (%d,%d,%dx%d)\n", S, p, x, y, w, h, want[p].x, want[p].y, want[p].w, want[p].h);
        bad++;
      }
      if (x % bs || y % bs || w % bs || h % bs || x % 4 || y % 4 || x + w > S || y + h > S) { printf("S %d part %d is not a union of %d-sample cells\n", S, p, bs); bad++; }
    }
  }
  if (bad) return 1;
  printf("ok\n");
  return 0;
}
