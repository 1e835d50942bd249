// kernels_staged.cuh -- gather kernels whose neighbour columns are staged in shared memory by ASYNCHRONOUS copies.
//
// Why (profiles/r1_ncu_top_kernels.md, r1_divdamp_variants.md): the round-1 gather kernels move their bytes at 0.3-0.5 of
// the HBM roofline with `long_scoreboard` as the only stall -- every thread walks a serial chain index -> gather ->
// index -> gather ..., and every attempt to widen it with registers lost the occupancy it bought.  Here the chain is
// cut to TWO round trips per block without spending a register on data in flight:
//   1. one round trip for everything indexed by the block's own entities (connectivity rows, own columns);
//   2. ALL neighbour level pairs are then requested at once with cp.async (LDGSTS, 16 bytes each, through L1 so that
//      neighbours shared by the columns of a tile are fetched once) into per-thread shared-memory slots;
//   3. cp.async.wait_all, and the stencil arithmetic reads shared memory in the reference's slot order.
// A thread only ever reads the slots it filled itself, so no barrier is needed for them.  Slots cover the first NS
// neighbours of a list; longer lists (heptagon pairs etc.) finish with plain loads.  Arithmetic and its order are
// unchanged: results are bit-identical to the plain kernels (tests/test_parity_gpu.py::test_staged_gathers_bit_identical).
#pragma once

DI void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
DI void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// slot j of thread tid: conflict-free (consecutive threads, consecutive 16-byte words)
#define SLOT(base, j) ((base)[(size_t)(j) * nthr + tid])

// u tendency :958-1163, staged form of k_dt_edge: the 2*n columns of the nonlinear Coriolis sum (u and pv_edge at
// edgesOnEdge) travel by cp.async; the eight cellsOnEdge gathers and the own columns are plain loads issued in the same
// phase (the block is capped by shared memory at 3 per SM, so their registers are free).
template <int NS>
__global__ void __launch_bounds__(224, 3) k_dt_edge_s(const View V, const DynTendParams P) {
  extern __shared__ __align__(16) double sm[];
  PAIR_THREAD(V.nEdges)
  const int TS = LP + 2;
  const int nthr = blockDim.x * blockDim.y, tid = threadIdx.y * blockDim.x + threadIdx.x;
  double* s_wduz = sm + (size_t)threadIdx.y * TS;
  D2* s_u = reinterpret_cast<D2*>(sm + (((size_t)blockDim.y * TS + 1) & ~(size_t)1));     // [NS][nthr]
  D2* s_pv = s_u + (size_t)NS * nthr;                                                      // [NS][nthr]
  const double* __restrict__ u = FLD(u); const double* __restrict__ pv = FLD(pv_edge);
  const int ME2 = V.maxEdges2;
  int4 cv = make_int4(0, 0, 0, 0);
  int n = 0;
  D2 u2 = bc(0.0), wduz = bc(0.0), rho_e = bc(0.0), pv_k = bc(0.0), tue = bc(0.0), trp = bc(0.0);
  D2 ke1 = bc(0.0), ke2 = bc(0.0), hd1 = bc(0.0), hd2 = bc(0.0), w1 = bc(0.0), w2 = bc(0.0);
  double w1n = 0.0, w2n = 0.0, invDc = 0.0;
  if (m0) {
    cv = V.ecv[x];
    n = V.nEdgesOnEdge[x];
    const int ns = n < NS ? n : NS;
#pragma unroll
    for (int j = 0; j < NS; ++j)
      if (j < ns) {
        const int eoe = V.edgesOnEdge[x * ME2 + j];
        cp_async16(&SLOT(s_u, j), u + (size_t)eoe * LP + k0);
        cp_async16(&SLOT(s_pv, j), pv + (size_t)eoe * LP + k0);
      }
    u2 = ld2(u, ix);
    const double* rw = FLD(rw);
    const D2 rw1 = G2(rw, cv.x), rw2 = G2(rw, cv.y);
    const double* ke = FLD(ke); const double* hd = FLD(h_divergence); const double* w = FLD(w);
    ke1 = G2(ke, cv.x); ke2 = G2(ke, cv.y); hd1 = G2(hd, cv.x); hd2 = G2(hd, cv.y); w1 = G2(w, cv.x); w2 = G2(w, cv.y);
    w1n = m1 ? G1(w, cv.x, k0 + 2) : 0.0; w2n = m1 ? G1(w, cv.y, k0 + 2) : 0.0;
    rho_e = ld2(FLD(rho_edge), ix); pv_k = ld2(pv, ix); tue = ld2(FLD(tend_u_euler), ix); trp = ld2(FLD(tend_ru_physics), ix);
    invDc = V.invDcEdge[x];
    const D2 rwavg = 0.5 * (rw1 + rw2);
    const D2 fzm = ld2(FLD(fzm), k0), fzp = ld2(FLD(fzp), k0);
    const D2 um = (k0 >= 2) ? ld2(u, ix - 2) : bc(0.0);            // (u[k0-2], u[k0-1])
    const double up = (k0 + 2 <= L) ? u[ix + 2] : 0.0;             // u[k1+1]
    wduz.x = wduz_at(k0, L, rwavg.x, fzm.x, fzp.x, um.x, um.y, u2.x, u2.y);
    if (m1) wduz.y = wduz_at(k1, L, rwavg.y, fzm.y, fzp.y, um.y, u2.x, u2.y, up);
    s_wduz[k0] = wduz.x; if (m1) s_wduz[k1] = wduz.y;
  }
  if (inx && k0 == L) s_wduz[L] = FLD(wduz)[ix];          // level L: never written, read as stored
  if (inx && k1 == L) s_wduz[L] = FLD(wduz)[ix + 1];
  cp_async_wait_all();
  __syncthreads();
  if (!m0) return;
  st2m(FLD(wduz), ix, wduz, m0, m1);
  const D2 wduz_p = mk(s_wduz[k1], m1 ? s_wduz[k1 + 1] : 0.0);
  D2 tend_u = -ld2(FLD(rdzw), k0) * (wduz_p - wduz);                                                // :987
  // nonlinear Coriolis term :991-1001 (Q14: nVertLevels * term, as in k_dt_edge)
  D2 q = bc(0.0);
  {
    const double Ld = (double)L;
    const int ns = n < NS ? n : NS;
#pragma unroll
    for (int j = 0; j < NS; ++j)
      if (j < ns) {
        const D2 workpv = 0.5 * (pv_k + SLOT(s_pv, j));
        q += Ld * (V.weightsOnEdge[x * ME2 + j] * SLOT(s_u, j) * workpv);
      }
    for (int j = NS; j < n; ++j) {                                   // lists longer than the staged slots
      const int eoe = V.edgesOnEdge[x * ME2 + j];
      const D2 workpv = 0.5 * (pv_k + G2(pv, eoe));
      q += Ld * (V.weightsOnEdge[x * ME2 + j] * G2(u, eoe) * workpv);
    }
  }
  st2m(FLD(q), ix, q, m0, m1);
  tend_u += rho_e * (q - (ke2 - ke1) * invDc) - u2 * 0.5 * (hd1 + hd2);                             // :1005-1007
  {
    const D2 w1p = mk(w1.y, w1n), w2p = mk(w2.y, w2n);
    const D2 wsum = w1 + w1p + w2 + w2p;
    tend_u -= (P.omega2 * V.cosAngleEdge[x] * V.cosLatEdge[x] * rho_e * 0.25 * wsum)
              - (u2 * 0.25 * wsum * rho_e * P.inv_r_earth);                                        // :1011-1017
  }
  if (P.rayleigh_u) {                                                                               // :1152-1159
    const int lim = L - P.rayleigh_levels + 1;
    if (k0 > lim) tend_u.x -= rho_e.x * u2.x * ((double)((double)k0 - (L - P.rayleigh_levels)) * P.rayleigh_coef_inverse);
    if (k1 > lim) tend_u.y -= rho_e.y * u2.y * ((double)((double)k1 - (L - P.rayleigh_levels)) * P.rayleigh_coef_inverse);
  }
  tend_u += tue + trp;                                                                              // :1162
  st2m(FLD(tend_u), ix, tend_u, m0, m1);
}

// staged form of k_acoustic_gather (:1644-1652): per slot i the edge's ru_p column and the theta_m column of the cell
// across the edge (the cell's own theta_m pair is a register)
template <int NS>
__global__ void __launch_bounds__(256, 4) k_acoustic_gather_s(const View V, double dts) {
  extern __shared__ __align__(16) double sm[];
  PAIR_THREAD_R()
  if (!m0) return;
  if (V.specZoneMaskCell[x] != 0.0) return;
  const int nthr = blockDim.x * blockDim.y, tid = threadIdx.y * blockDim.x + threadIdx.x;
  D2* s_ru = reinterpret_cast<D2*>(sm);            // [NS][nthr]
  D2* s_tm = s_ru + (size_t)NS * nthr;             // [NS][nthr]  theta_m of the neighbour across slot i
  const int ME = V.maxEdges, n = V.nEdgesOnCell[x];
  const double* __restrict__ ru_p = FLD(ru_p); const double* __restrict__ tm = FLD(theta_m);
  const int ns = n < NS ? n : NS;
  int c1s[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    c1s[i] = x;
    if (i < ns) {
      const int e = V.edgesOnCell[x * V.MEP + i];
      const int c1 = V.c1OnCell[x * V.MEP + i], c2 = V.c2OnCell[x * V.MEP + i];
      c1s[i] = c1;
      cp_async16(&SLOT(s_ru, i), ru_p + (size_t)e * LP + k0);
      // exactly one of (c1, c2) is this cell on a well-formed mesh; if neither is, c2 keeps a plain load below
      cp_async16(&SLOT(s_tm, i), tm + (size_t)(c1 == x ? c2 : c1) * LP + k0);
    }
  }
  const D2 tm_own = ld2(tm, ix);
  const double inva = V.invAreaCell[x];
  cp_async_wait_all();
  D2 rs = bc(0), ts = bc(0);
#pragma unroll
  for (int i = 0; i < NS; ++i)
    if (i < ns) {
      const int c2 = V.c2OnCell[x * V.MEP + i];
      const D2 flux = V.edgesOnCellSign[x * ME + i] * dts * V.dvOnCell[x * ME + i] * SLOT(s_ru, i) * inva;
      rs -= flux;
      const bool own1 = c1s[i] == x;
      const D2 t1 = own1 ? tm_own : SLOT(s_tm, i);                         // theta_m(c1)
      const D2 t2 = own1 ? SLOT(s_tm, i) : (c2 == x ? tm_own : G2(tm, c2)); // theta_m(c2)
      ts -= flux * 0.5 * (t2 + t1);
    }
  for (int i = NS; i < n; ++i) {
    const int e = V.edgesOnCell[x * V.MEP + i];
    const int c1 = V.c1OnCell[x * V.MEP + i], c2 = V.c2OnCell[x * V.MEP + i];
    const D2 flux = V.edgesOnCellSign[x * ME + i] * dts * V.dvOnCell[x * ME + i] * G2(ru_p, e) * inva;
    rs -= flux;
    ts -= flux * 0.5 * (G2(tm, c2) + G2(tm, c1));
  }
  st2m(V.scr_rs, ix, rs, m0, m1); st2m(V.scr_ts, ix, ts, m0, m1);
}

// staged form of k_dt_theta_flux (:1333-1340): the advection stencil's theta_m columns
template <int NS>
__global__ void __launch_bounds__(256, 5) k_dt_theta_flux_s(const View V) {
  extern __shared__ __align__(16) double sm[];
  PAIR_THREAD(V.nEdges)
  if (!m0) return;
  const int nthr = blockDim.x * blockDim.y, tid = threadIdx.y * blockDim.x + threadIdx.x;
  D2* s_tm = reinterpret_cast<D2*>(sm);            // [NS][nthr]
  const int NA = V.nAdv;
  const double* __restrict__ tm = FLD(theta_m);
  const int na = V.nAdvCellsForEdge[x];
  const int ns = na < NS ? na : NS;
#pragma unroll
  for (int j = 0; j < NS; ++j)
    if (j < ns) cp_async16(&SLOT(s_tm, j), tm + (size_t)V.advCellsForEdge[x * NA + j] * LP + k0);
  const D2 sg = sgn1(ld2(FLD(ru), ix));
  cp_async_wait_all();
  D2 fa = bc(0.0);
#pragma unroll
  for (int j = 0; j < NS; ++j)
    if (j < ns) {
      const D2 sw = V.adv_coefs[x * NA + j] + sg * V.adv_coefs_3rd[x * NA + j];
      fa += sw * SLOT(s_tm, j);
    }
  for (int j = NS; j < na; ++j) {
    const D2 sw = V.adv_coefs[x * NA + j] + sg * V.adv_coefs_3rd[x * NA + j];
    fa += sw * G2(tm, V.advCellsForEdge[x * NA + j]);
  }
  st2m(V.scr_flux, ix, fa, m0, m1);
}
