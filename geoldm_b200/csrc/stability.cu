// Bond-order stability metric (qm9/analyze.py:209-245, qm9/bond_analyze.py:101-146) for a ragged batch.
// One warp per molecule; lane a owns atoms a, a+32, ...; every atom walks over all partners (n <= a few hundred, the
// coordinates of one molecule stay in L1).  Arithmetic follows the reference's float32 numpy expression exactly:
//   d = 100 * sqrt((dx*dx + dy*dy) + dz*dz)   (no FMA contraction), thresholds are small integers held as fp32.
#include "common.cuh"
#include "../../include/geoldm_b200.h"

namespace {

__device__ __forceinline__ int bond_order(float d100, const float* __restrict__ thr, int T, int ta, int tb) {
  const int idx = ta * T + tb;
  const float t1 = thr[idx];
  if (t1 < 0.f || !(d100 < t1)) return 0;
  const float t2 = thr[T * T + idx];
  if (t2 >= 0.f && d100 < t2) {
    const float t3 = thr[2 * T * T + idx];
    if (t3 >= 0.f && d100 < t3) return 3;
    return 2;
  }
  return 1;
}

__global__ void stability_kernel(int n_mol, const int* __restrict__ mol_off, const float* __restrict__ x,
                                 const int* __restrict__ atom_type, int T, const float* __restrict__ thr,
                                 const int* __restrict__ allowed, int sorted_pair, int* __restrict__ nr_bonds,
                                 int* __restrict__ n_stable) {
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= n_mol) return;
  const int n0 = mol_off[m], n = mol_off[m + 1] - n0;
  int stable = 0;
  for (int a = lane; a < n; a += 32) {
    const float ax = x[3 * (size_t)(n0 + a)], ay = x[3 * (size_t)(n0 + a) + 1], az = x[3 * (size_t)(n0 + a) + 2];
    const int ta = atom_type[n0 + a];
    int bonds = 0;
    for (int b = 0; b < n; ++b) {
      if (b == a) continue;
      const float* p = x + 3 * (size_t)(n0 + b);
      const float dx = __fsub_rn(ax, p[0]), dy = __fsub_rn(ay, p[1]), dz = __fsub_rn(az, p[2]);
      const float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      const float d100 = __fmul_rn(100.f, __fsqrt_rn(s));
      const int tb = atom_type[n0 + b];
      int t1, t2;                                    // ordered lookup: first atom of the pair, second atom
      if (sorted_pair) { t1 = min(ta, tb); t2 = max(ta, tb); }
      else if (a < b)  { t1 = ta; t2 = tb; }
      else             { t1 = tb; t2 = ta; }
      bonds += bond_order(d100, thr, T, t1, t2);
    }
    if (nr_bonds) nr_bonds[n0 + a] = bonds;
    stable += (bonds < 32 && ((unsigned)allowed[ta] >> bonds) & 1u) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) stable += __shfl_xor_sync(0xffffffffu, stable, o);
  if (lane == 0) n_stable[m] = stable;
}

}  // namespace

extern "C" int geoldm_stability(int n_mol, const int* mol_off, const float* x, const int* atom_type, int n_types,
                                const float* thr, const int* allowed, int sorted_pair, int* nr_bonds, int* n_stable,
                                void* stream) {
  GEOLDM_REQUIRE(n_mol >= 0 && n_types > 0 && n_types <= 64, "stability: n_mol %d n_types %d", n_mol, n_types);
  GEOLDM_REQUIRE(mol_off && x && atom_type && thr && allowed && n_stable, "stability: null argument%s", "");
  if (n_mol == 0) return 0;
  stability_kernel<<<(n_mol * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_mol, mol_off, x, atom_type, n_types, thr,
                                                                              allowed, sorted_pair, nr_bonds, n_stable);
  GEOLDM_CHECK_LAUNCH("stability_kernel");
  return 0;
}
