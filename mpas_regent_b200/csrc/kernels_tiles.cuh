// kernels_tiles.cuh -- tile-staged forms of the horizontal gather kernels (MpasConfig.edge_tiles).
//
// The plain gather kernels read every neighbour column with per-thread 128-bit loads after an index load; their
// measured limit is the latency of those dependent loads together with the L1 data pipe.  Here a
// block owns a TILE of TE consecutive entities; the set of DISTINCT neighbour columns the whole tile touches is
// resolved once on the host at upload_mesh (static connectivity; lists in internal numbering), and the block fetches
// each distinct column exactly once into shared memory with cp.async.bulk strips (LP*8 contiguous bytes per column
// and field; completion on one mbarrier).  The bytes in flight cost no registers and do not pass through L1; the
// arithmetic then reads shared memory in the reference's slot order, so results are bit-identical to the plain kernels.
// A neighbour that does not fit in the tile's CAP slots keeps its global load (slot id 255).
#pragma once
#include "kernels.cuh"


DI int slot_byte(const uint4& a, int j) {      // byte j (0..15) of a 16-byte row held in four registers
  const unsigned w = (j < 8) ? ((j < 4) ? a.x : a.y) : ((j < 12) ? a.z : a.w);
  return (int)((w >> ((j & 3) * 8)) & 255u);
}

// u tendency  :958-1163 -- same arithmetic, expression by expression, as k_dt_edge (kernels.cuh)
// MAXT = LP/2 * TE threads at most (checked by the host), MINB resident blocks asked from ptxas
// ABL: ablation switches for profiling only (1 = no bulk copies and no wait: consumer side alone; 2 = copies + wait, then one store: staging alone)
template <int TE, int MAXT, int MINB, int ABL = 0>
__global__ void __launch_bounds__(MAXT, MINB) k_dt_edge_tile(const View V, const DynTendParams P, const EdgeTiles T) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int k0 = 2 * (int)threadIdx.x, k1 = k0 + 1;
  const int LP = V.LP, L = V.L;
  const int tile = blockIdx.x;
  const int x = tile * TE + (int)threadIdx.y;
  const bool inx = x < V.nEdges;
  const size_t ix = (size_t)(inx ? x : 0) * LP + k0;
  const bool m0 = inx && k0 < L, m1 = inx && k1 < L;
  const int ME2 = V.maxEdges2;
  // shared memory: [cap][2][LP] staged columns (u, pv_edge) | [TE][ME2] weights | [TE][LP+2] wduz | mbarrier
  double* s_col = reinterpret_cast<double*>(smraw);
  double* s_wt = s_col + (size_t)T.cap * 2 * LP;
  const int TS = LP + 2;
  double* s_wduz = s_wt + (size_t)TE * ME2 + (size_t)threadIdx.y * TS;
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_wt + (size_t)TE * ME2 + (size_t)TE * TS);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
  const double* u = FLD(u);
  const double* pv = FLD(pv_edge);
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  if (!(ABL & 1)) {
    const int nc = T.ncols[tile];
    const int nrow = min(TE, V.nEdges - tile * TE);
    const uint32_t colB = (uint32_t)(LP * sizeof(double)), wtB = (uint32_t)(nrow * ME2 * sizeof(double));
    if (tid == 0) mbar_expect_tx(bar, (uint32_t)(2 * nc) * colB + wtB);
    const int* cl = T.cols + (size_t)tile * T.cap;
    for (int i = tid; i < 2 * nc; i += nthr) {
      const int s = i >> 1, f = i & 1;
      bulk_g2s(s_col + ((size_t)s * 2 + f) * LP, (f ? pv : u) + (size_t)cl[s] * LP, colB, bar);
    }
    if (tid == nthr - 1) bulk_g2s(s_wt, V.weightsOnEdge + (size_t)tile * TE * ME2, wtB, bar);
  }
  // loads that do not depend on the staged columns are issued while the strips are in flight
  int4 cv = make_int4(0, 0, 0, 0);
  uint4 sa = make_uint4(~0u, ~0u, ~0u, ~0u);
  D2 rwavg = bc(0.0), rho_e = bc(0.0);
  int n = 0;
  if (m0) {
    cv = V.ecv[x];
    n = V.nEdgesOnEdge[x];
    sa = *reinterpret_cast<const uint4*>(T.slot + (size_t)x * T.SP);
    const double* rw = FLD(rw);
    rwavg = 0.5 * (G2(rw, cv.x) + G2(rw, cv.y));
    rho_e = ld2(FLD(rho_edge), ix);
  }
  if (!(ABL & 1)) mbar_wait(bar, 0);
  if (ABL & 2) { if (m0) st2m(FLD(tend_u), ix, ld2(s_col + (size_t)threadIdx.y * 2 * LP, k0) + rwavg + rho_e, m0, m1); return; }
  const double* s_u = s_col + (size_t)threadIdx.y * 2 * LP;      // own column: slot = local index
  const double* s_pv = s_u + LP;
  D2 u2 = bc(0.0), wduz = bc(0.0);
  if (m0) {
    u2 = ld2(s_u, k0);
    const D2 fzm = ld2(FLD(fzm), k0), fzp = ld2(FLD(fzp), k0);
    const D2 um = (k0 >= 2) ? ld2(s_u, k0 - 2) : bc(0.0);            // (u[k0-2], u[k0-1])
    const double up = (k0 + 2 <= L) ? s_u[k0 + 2] : 0.0;             // u[k1+1]
    wduz.x = wduz_at(k0, L, rwavg.x, fzm.x, fzp.x, um.x, um.y, u2.x, u2.y);
    if (m1) wduz.y = wduz_at(k1, L, rwavg.y, fzm.y, fzp.y, um.y, u2.x, u2.y, up);
    s_wduz[k0] = wduz.x; if (m1) s_wduz[k1] = wduz.y;
    st2m(FLD(wduz), ix, wduz, m0, m1);
  }
  if (inx && k0 == L) s_wduz[L] = FLD(wduz)[ix];          // level L: never written, read as stored
  if (inx && k1 == L) s_wduz[L] = FLD(wduz)[ix + 1];
  __syncthreads();
  if (!m0) return;
  const D2 wduz_p = mk(s_wduz[k1], m1 ? s_wduz[k1 + 1] : 0.0);
  D2 tend_u = -ld2(FLD(rdzw), k0) * (wduz_p - wduz);                                                              // :987
  // nonlinear Coriolis term :991-1001 (Q14: added once, multiplied by nVertLevels, as in k_dt_edge)
  D2 q = bc(0.0);
  {
    const D2 pv_k = ld2(s_pv, k0);
    const double Ld = (double)L;
    const double* wrow = s_wt + (size_t)threadIdx.y * ME2;
    auto term = [&](int j, double wt) {
      const int s = (j < 16) ? slot_byte(sa, j) : (int)T.slot[(size_t)x * T.SP + j];
      D2 pvj, uj;
      if (s != 255) {
        const double* c = s_col + (size_t)s * 2 * LP + k0;
        uj = ld2(c, 0); pvj = ld2(c, LP);
      } else {
        const int eoe = V.edgesOnEdge[(size_t)x * ME2 + j];
        uj = G2(u, eoe); pvj = G2(pv, eoe);
      }
      const D2 workpv = 0.5 * (pv_k + pvj);
      q += Ld * (wt * uj * workpv);
    };
    int j0 = 0;
#pragma unroll 1
    for (; j0 + 1 < n; j0 += 2) {                // the host enables this kernel only when maxEdges2 is even: weight rows are 16-byte aligned
      const D2 wt = ld2(wrow, j0);
      term(j0, wt.x);
      term(j0 + 1, wt.y);
    }
    if (j0 < n) term(j0, wrow[j0]);
  }
  st2m(FLD(q), ix, q, m0, m1);
  const double* ke = FLD(ke); const double* hd = FLD(h_divergence); const double* w = FLD(w);
  tend_u += rho_e * (q - (G2(ke, cv.y) - G2(ke, cv.x)) * V.invDcEdge[x]) - u2 * 0.5 * (G2(hd, cv.x) + G2(hd, cv.y));   // :1005-1007
  {
    const D2 w1 = G2(w, cv.x), w2 = G2(w, cv.y);
    const D2 w1p = mk(w1.y, m1 ? G1(w, cv.x, k0 + 2) : 0.0), w2p = mk(w2.y, m1 ? G1(w, cv.y, k0 + 2) : 0.0);
    const D2 wsum = w1 + w1p + w2 + w2p;
    tend_u -= (P.omega2 * V.cosAngleEdge[x] * V.cosLatEdge[x] * rho_e * 0.25 * wsum)
              - (u2 * 0.25 * wsum * rho_e * P.inv_r_earth);                                        // :1011-1017
  }
  if (P.rayleigh_u) {                                                                               // :1152-1159
    const int lim = L - P.rayleigh_levels + 1;
    if (k0 > lim) tend_u.x -= rho_e.x * u2.x * ((double)((double)k0 - (L - P.rayleigh_levels)) * P.rayleigh_coef_inverse);
    if (k1 > lim) tend_u.y -= rho_e.y * u2.y * ((double)((double)k1 - (L - P.rayleigh_levels)) * P.rayleigh_coef_inverse);
  }
  tend_u += ld2(FLD(tend_u_euler), ix) + ld2(FLD(tend_ru_physics), ix);                                                                             // :1162
  st2m(FLD(tend_u), ix, tend_u, m0, m1);
}
