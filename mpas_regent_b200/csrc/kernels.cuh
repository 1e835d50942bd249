// kernels.cuh -- hand-written sm_100a kernels for the RK3 dynamics hot path.
//
// Thread mapping (all stencil kernels): blockDim = (LP, CPB): threadIdx.x = level k (levels
// contiguous in memory -> every warp request is a run of consecutive doubles of one or two
// columns), threadIdx.y = column within the block, CPB chosen so LP*CPB is a multiple of 32.
// A block therefore owns CPB whole columns, so column neighbours (k-2..k+1) computed in the
// same kernel are exchanged through shared memory; neighbour columns (edgesOnCell,
// cellsOnEdge, advCellsForEdge ...) are gathered straight from L2/HBM as contiguous level
// strips.  Flux divergences are cell-centric gathers over edgesOnCell in slot order: no
// atomics, bit-reproducible.  Every kernel is HBM-bound (fp64, ~0.3 flop/B).
//
// Arithmetic follows the reference's expression order (compiled with --fmad=false) so that
// results agree with the CPU oracle to the last bit wherever no libm call is involved.
// Each kernel cites the reference lines it implements (dynamics/dynamics_tasks.rg).
#pragma once
#include "view.h"

#define FLD(name) (V.f[MPASB200_F_##name])
#define COLUMN_THREAD(n)                                        \
  const int k = threadIdx.x;                                    \
  const int x = blockIdx.x * blockDim.y + threadIdx.y;          \
  const bool inx = x < (n);                                     \
  const int LP = V.LP; const int L = V.L;                       \
  const size_t ix = (size_t)(inx ? x : 0) * LP + k;             \
  (void)L; (void)ix;
#define AT(p, x_, k_) ((p)[(size_t)(x_) * LP + (k_)])

__device__ __forceinline__ double flux4(double q_im2, double q_im1, double q_i, double q_ip1, double ua) {
  return ua * (7. * (q_i + q_im1) - (q_ip1 + q_im2)) / 12.0;                       // :781-783
}
__device__ __forceinline__ double flux3(double q_im2, double q_im1, double q_i, double q_ip1, double ua, double coef3) {
  return flux4(q_im2, q_im1, q_i, q_ip1, ua) + coef3 * fabs(ua) * ((q_ip1 - q_im2) - 3. * (q_i - q_im1)) / 12.0;   // :786-789
}
__device__ __forceinline__ double dmin(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double dmax(double a, double b) { return (a < b) ? b : a; }

// ============================================================================================
// atm_rk_integration_setup  :747-778
__global__ void k_setup_cell(const View V) {
  COLUMN_THREAD(V.nCells)
  if (!inx || k >= L) return;
  FLD(rw_save)[ix] = FLD(rw)[ix]; FLD(rtheta_p_save)[ix] = FLD(rtheta_p)[ix]; FLD(rho_p_save)[ix] = FLD(rho_p)[ix];
  FLD(w_2)[ix] = FLD(w)[ix]; FLD(theta_m_2)[ix] = FLD(theta_m)[ix];
  const double r = FLD(rho_zz)[ix]; FLD(rho_zz_2)[ix] = r; FLD(rho_zz_old_split)[ix] = r;
}
__global__ void k_setup_edge(const View V) {
  COLUMN_THREAD(V.nEdges)
  if (!inx || k >= L) return;
  FLD(ru_save)[ix] = FLD(ru)[ix]; FLD(u_2)[ix] = FLD(u)[ix];
}

// atm_compute_moist_coefficients  :460-502   (qtot = 0; cqw from it for k > 0; cqu never written)
__global__ void k_moist(const View V) {
  COLUMN_THREAD(V.nCells)
  if (!inx || k >= L) return;
  const double qz = 0.0;
  FLD(qtot)[ix] = qz;
  if (k > 0) { const double qtotal = 0.5 * (qz + qz); FLD(cqw)[ix] = 1.0 / (1.0 + qtotal); }
}

// atm_compute_vert_imp_coefs  :513-592
// One block owns whole columns: everything read from the previous call (gamma_tri[k-1]) is read
// before the barrier, everything this call produces for neighbours (coftz, cofwt) goes through smem.
__global__ void k_vert_imp(const View V, double dtseps, double c2, double rcv, double gravity) {
  extern __shared__ double sm[];
  COLUMN_THREAD(V.nCells)
  const int TS = LP + 1;
  double* s_coftz = sm + (size_t)threadIdx.y * TS;
  double* s_cofwt = sm + (size_t)(blockDim.y + threadIdx.y) * TS;
  const double* rdzw = FLD(rdzw); const double* rdzu = FLD(rdzu); const double* fzm = FLD(fzm); const double* fzp = FLD(fzp);
  const bool act = inx && k < L;
  double zz_k = 0, zz_m = 0, coftz_k = 0, cofwt_k = 0, cofwr_k = 0, cofwz_k = 0, gamma_prev = 0;
  if (blockIdx.x == 0 && threadIdx.y == 0 && k < L) FLD(cofrz)[k] = dtseps * rdzw[k];          // :537-539
  if (act) {
    zz_k = FLD(zz)[ix];
    const double ex_k = FLD(exner)[ix], th_k = FLD(theta_m)[ix];
    if (k > 0) {
      zz_m = FLD(zz)[ix - 1];
      const double zf = fzm[k] * zz_k + fzp[k] * zz_m;
      cofwr_k = .5 * dtseps * gravity * zf;                                                      // :552
      cofwz_k = dtseps * c2 * zf * rdzu[k] * FLD(cqw)[ix] * (fzm[k] * ex_k + fzp[k] * FLD(exner)[ix - 1]);   // :557
      coftz_k = dtseps * (fzm[k] * th_k + fzp[k] * FLD(theta_m)[ix - 1]);                        // :558
      gamma_prev = (k == 1) ? 0.0 : FLD(gamma_tri)[ix - 1];                                      // :545, :583 (previous call's gamma)
    }
    const double qtotal = FLD(qtot)[ix];
    cofwt_k = .5 * dtseps * rcv * zz_k * gravity * FLD(rho_base)[ix] / (1.0 + qtotal) * ex_k
              / ((FLD(rtheta_base)[ix] + FLD(rtheta_p)[ix]) * FLD(exner_base)[ix]);             // :563
    s_coftz[k] = coftz_k; s_cofwt[k] = cofwt_k;
  }
  if (inx && k == L) s_coftz[L] = FLD(coftz)[ix];     // level L is never written: whatever the mirror holds
  __syncthreads();
  if (!act) return;
  FLD(coftz)[ix] = coftz_k; FLD(cofwt)[ix] = cofwt_k;
  if (k == 0) { FLD(gamma_tri)[ix] = 0.0; return; }
  FLD(cofwr)[ix] = cofwr_k; FLD(cofwz)[ix] = cofwz_k;
  const double coftz_m = s_coftz[k - 1], coftz_p = s_coftz[k + 1], cofwt_m = s_cofwt[k - 1];
  const double cofrz_k = dtseps * rdzw[k], cofrz_m = dtseps * rdzw[k - 1];
  const double a = -1.0 * cofwz_k * coftz_m * rdzw[k - 1] * zz_m + cofwr_k * cofrz_m - cofwt_m * coftz_m * rdzw[k - 1];   // :568-569
  const double b = 1.0 + cofwz_k * (coftz_k * rdzw[k] * zz_k + coftz_k * rdzw[k - 1] * zz_m)
                   - coftz_k * (cofwt_k * rdzw[k] - cofwt_k * rdzw[k - 1]) + cofwr_k * ((cofrz_k - cofrz_m));           // :571-573
  const double c = -1.0 * cofwz_k * coftz_p * rdzw[k] * zz_k - cofwr_k * cofrz_k + cofwt_k * coftz_p * rdzw[k];          // :575-576
  const double alpha = 1.0 / (b - a * gamma_prev);                                                                      // :583
  FLD(a_tri)[ix] = a; FLD(b_tri)[ix] = b; FLD(c_tri)[ix] = c; FLD(alpha_tri)[ix] = alpha;
  FLD(gamma_tri)[ix] = c * alpha;                                                                                       // :589
}

// ============================================================================================
// atm_compute_solve_diagnostics  :328-454
__global__ void k_diag_vertex(const View V) {      // vorticity :356-366, pv_vertex :443-445
  COLUMN_THREAD(V.nVertices)
  if (!inx || k >= L) return;
  const int VD = V.vertexDegree;
  const double* u = FLD(u);
  double vort = 0.0;
  for (int i = 0; i < VD; ++i) {
    const int e = V.edgesOnVertex[x * VD + i];
    const double s = V.edgesOnVertexSign[x * VD + i] * V.dcEdge[e];
    vort += s * AT(u, e, k);
  }
  vort *= V.invAreaTriangle[x];
  FLD(vorticity)[ix] = vort;
  FLD(pv_vertex)[ix] = V.fVertex[x] + vort;
}
__global__ void k_diag_cell(const View V) {        // divergence :369-379 (s + u), ke :382-390
  COLUMN_THREAD(V.nCells)
  if (!inx || k >= L) return;
  const int ME = V.maxEdges, n = V.nEdgesOnCell[x];
  const double* u = FLD(u);
  double div = 0.0, kec = 0.0;
  for (int i = 0; i < n; ++i) {
    const int e = V.edgesOnCell[x * ME + i];
    const double s = V.edgesOnCellSign[x * ME + i] * V.dvEdge[e];
    const double ue = AT(u, e, k);
    div += s + ue;
    const double efac = V.dcEdge[e] * V.dvEdge[e];
    kec += 0.25 * (efac * (ue * ue));               // ke_edge recomputed from u: same value as the stored field
  }
  const double r = V.invAreaCell[x];
  FLD(divergence)[ix] = div * r;
  FLD(ke)[ix] = kec * r;
}
template <bool RECON_V>
__global__ void k_diag_edge(const View V) {        // h_edge, ke_edge :346-353; v :431-438; pv_edge :449-451
  COLUMN_THREAD(V.nEdges)
  if (!inx || k >= L) return;
  const int c1 = V.cellsOnEdge[x * 2], c2 = V.cellsOnEdge[x * 2 + 1];
  const double* u = FLD(u); const double* h = FLD(h); const double* pvv = FLD(pv_vertex);
  FLD(h_edge)[ix] = 0.5 * (AT(h, c1, k) + AT(h, c2, k));
  const double efac = V.dcEdge[x] * V.dvEdge[x];
  const double ue = u[ix];
  FLD(ke_edge)[ix] = efac * (ue * ue);
  if (RECON_V) {
    const int ME2 = V.maxEdges2, n = V.nEdgesOnEdge[x];
    double vv = 0;
    for (int i = 1; i < n; ++i) {                   // starts at 1 (Q7)
      const int eoe = V.edgesOnEdge_ECP[x * ME2 + i];
      vv += V.weightsOnEdge[x * ME2 + i] * AT(u, eoe, k);
    }
    FLD(v)[ix] = vv;
  }
  FLD(pv_edge)[ix] = 0.5 * (AT(pvv, V.verticesOnEdge[x * 2], k) + AT(pvv, V.verticesOnEdge[x * 2 + 1], k));
}
__global__ void k_diag_ke_vertex(const View V) {   // hollingsworth part 1 :395-400
  COLUMN_THREAD(V.nVertices)
  if (!inx || k >= L) return;
  const int VD = V.vertexDegree;
  const double* ke_edge = FLD(ke_edge);
  const double r = 0.25 * V.invAreaTriangle[x];
  FLD(ke_vertex)[ix] = (AT(ke_edge, V.edgesOnVertex[x * VD], k) + AT(ke_edge, V.edgesOnVertex[x * VD + 1], k)
                        + AT(ke_edge, V.edgesOnVertex[x * VD + 2], k)) * r;
}
__global__ void k_diag_ke_holl(const View V) {     // hollingsworth part 2 :403-417
  COLUMN_THREAD(V.nCells)
  if (!inx || k >= L) return;
  const int ME = V.maxEdges, VD = V.vertexDegree, n = V.nEdgesOnCell[x];
  const double ke_fact = 1.0 - 0.375;
  const double* kev = FLD(ke_vertex);
  double kec = FLD(ke)[ix] * ke_fact;
  const double r = V.invAreaCell[x];
  for (int i = 0; i < n; ++i) {
    const int iv = V.verticesOnCell[x * ME + i];
    const int j = V.kiteForCell[x * ME + i];
    kec += (1.0 - ke_fact) * V.kiteAreasOnVertex[iv * VD + j] * AT(kev, iv, k) * r;
  }
  FLD(ke)[ix] = kec;
}

// ============================================================================================
// atm_compute_dyn_tend_work  :814-1480
// cell pre-pass: kdiff (:858-917), h_divergence (:924-938), tend_rho + dpdz (:942-951)
template <bool RK0>
__global__ void k_dt_cell0(const View V, const DynTendParams P, double len_disp, double cam_coef) {
  COLUMN_THREAD(V.nCells)
  if (!inx || k >= L) return;
  const int ME = V.maxEdges, n = V.nEdgesOnCell[x];
  const double* ru = FLD(ru);
  double hdiv = 0.0;
  if (RK0 && P.mixing == MPASB200_MIX_2D_SMAGORINSKY) {
    const double* u = FLD(u); const double* v = FLD(v);
    double d_diag = 0.0, d_off = 0.0;
    for (int i = 0; i < n; ++i) {
      const int e = V.edgesOnCell[x * ME + i];
      const double a = V.defc_a[x * ME + i], b = V.defc_b[x * ME + i];
      const double ue = AT(u, e, k), ve = AT(v, e, k);
      d_diag += a * ue - b * ve;
      d_off += b * ue + a * ve;
      hdiv += (V.edgesOnCell_sign[x * ME + i] * V.dvEdge[e]) * AT(ru, e, k);
    }
    double kd = dmin(P.kdiff_scale * sqrt(d_diag * d_diag + d_off * d_off), P.kdiff_cap);      // :884-886
    if (P.cam_on && k >= L - 2) kd = dmax(kd, pow(2.0, (double)(k - (L - 2))) * 2.0833 * len_disp * cam_coef);   // :911-914
    FLD(kdiff)[ix] = kd;
  } else {
    for (int i = 0; i < n; ++i) {
      const int e = V.edgesOnCell[x * ME + i];
      hdiv += (V.edgesOnCell_sign[x * ME + i] * V.dvEdge[e]) * AT(ru, e, k);
    }
    if (RK0 && (P.mixing == MPASB200_MIX_2D_FIXED || P.cam_on)) {
      double kd = (P.mixing == MPASB200_MIX_2D_FIXED) ? 0.0 : FLD(kdiff)[ix];
      if (P.cam_on && k >= L - 2) kd = dmax(kd, pow(2.0, (double)(k - (L - 2))) * 2.0833 * len_disp * cam_coef);
      FLD(kdiff)[ix] = kd;
    }
  }
  hdiv *= V.invAreaCell[x];
  FLD(h_divergence)[ix] = hdiv;
  if (RK0) {
    const double* rw = FLD(rw);
    const double qt = FLD(qtot)[ix];
    FLD(tend_rho)[ix] = -hdiv - FLD(rdzw)[k] * (rw[ix + 1] - rw[ix] + FLD(tend_rho_physics)[ix]);   // :947
    FLD(dpdz)[ix] = -P.gravity * (FLD(rho_base)[ix] * (qt) + FLD(rho_p_save)[ix] * (1.0 + qt));   // :949
  }
}

// first del^2 of u  :1030-1043  (only delsq_u; the tend_u_euler contribution is added in k_dt_edge)
__global__ void k_dt_edge_delsq(const View V) {
  COLUMN_THREAD(V.nEdges)
  if (!inx || k >= L) return;
  const int c1 = V.cellsOnEdge[x * 2], c2 = V.cellsOnEdge[x * 2 + 1];
  const int v1 = V.verticesOnEdge[x * 2], v2 = V.verticesOnEdge[x * 2 + 1];
  const double* dv = FLD(divergence); const double* vo = FLD(vorticity);
  const double r_dc = V.invDcEdge[x];
  const double r_dv = dmin(V.invDvEdge[x], 4 * r_dc);
  const double u_diffusion = (AT(dv, c2, k) - AT(dv, c1, k)) * r_dc - (AT(vo, v2, k) - AT(vo, v1, k)) * r_dv;
  FLD(delsq_u)[ix] = 0.0 + u_diffusion;
}
__global__ void k_dt_vertex_delsq(const View V) {   // delsq_vorticity :1052-1060
  COLUMN_THREAD(V.nVertices)
  if (!inx || k >= L) return;
  const int VD = V.vertexDegree;
  const double* dsu = FLD(delsq_u);
  double acc = 0.0;
  for (int i = 0; i < VD; ++i) {
    const int e = V.edgesOnVertex[x * VD + i];
    const double edge_sign = V.invAreaTriangle[x] * V.dcEdge[e] * V.edgesOnVertex_sign[x * VD + i];
    acc += edge_sign * AT(dsu, e, k);
  }
  FLD(delsq_vorticity)[ix] = acc;
}
__global__ void k_dt_cell_delsq(const View V) {     // delsq_divergence :1062-1070
  COLUMN_THREAD(V.nCells)
  if (!inx || k >= L) return;
  const int ME = V.maxEdges, n = V.nEdgesOnCell[x];
  const double* dsu = FLD(delsq_u);
  const double r = V.invAreaCell[x];
  double acc = 0.0;
  for (int i = 0; i < n; ++i) {
    const int e = V.edgesOnCell[x * ME + i];
    const double edge_sign = r * V.dvEdge[e] * V.edgesOnCell_sign[x * ME + i];
    acc += edge_sign * AT(dsu, e, k);
  }
  FLD(delsq_divergence)[ix] = acc;
}

// u tendency  :958-1163
template <bool RK0>
__global__ void k_dt_edge(const View V, const DynTendParams P) {
  extern __shared__ double sm[];
  COLUMN_THREAD(V.nEdges)
  const int TS = LP + 1;
  double* s_wduz = sm + (size_t)threadIdx.y * TS;
  double* s_um = sm + (size_t)(blockDim.y + threadIdx.y) * TS;      // u_mix (vertical mixing of the perturbation)
  const bool act = inx && k < L;
  const double* u = FLD(u);
  int c1 = 0, c2 = 0;
  double u_k = 0, wduz_k = 0.0;
  if (act) {
    c1 = V.cellsOnEdge[x * 2]; c2 = V.cellsOnEdge[x * 2 + 1];
    u_k = u[ix];
    const double* rw = FLD(rw);
    const double* fzm = FLD(fzm); const double* fzp = FLD(fzp);
    if (k == 1 || k == L - 1)                                                                       // :974-976
      wduz_k = 0.5 * (AT(rw, c1, k) + AT(rw, c2, k)) * (fzm[k] * u_k + fzp[k] * u[ix - 1]);
    if (k > 1 && k < L - 1)                                                                         // :977-980
      wduz_k = flux3(u[ix - 2], u[ix - 1], u_k, u[ix + 1], 0.5 * (AT(rw, c1, k) + AT(rw, c2, k)), 1.0);
    s_wduz[k] = wduz_k;
    FLD(wduz)[ix] = wduz_k;
    if (RK0 && P.vmix_u_on && !P.mix_full) {                                                        // :1120-1123
      const double um = u_k - FLD(u_init)[k] * V.cosAngleEdge[x] - FLD(v_init)[k] * V.sinAngleEdge[x];
      s_um[k] = um; FLD(u_mix)[ix] = um;
    }
  }
  if (inx && k == L) s_wduz[L] = FLD(wduz)[ix];          // level L: never written, read as stored
  __syncthreads();
  if (!act) return;
  const double rho_e = FLD(rho_edge)[ix];
  const double invDc = V.invDcEdge[x];
  double tend_u = -FLD(rdzw)[k] * (s_wduz[k + 1] - wduz_k);                                         // :987
  // nonlinear Coriolis term :991-1001.  The reference adds each term nVertLevels times in a row
  // (Q14); here it is added once, multiplied by nVertLevels (same value to O(L) ulp).
  double q = 0.0;
  {
    const int ME2 = V.maxEdges2, n = V.nEdgesOnEdge[x];
    const double* pv = FLD(pv_edge);
    const double pv_k = pv[ix];
    const double Ld = (double)L;
    for (int j = 0; j < n; ++j) {
      const int eoe = V.edgesOnEdge[x * ME2 + j];
      const double workpv = 0.5 * (pv_k + AT(pv, eoe, k));
      q += Ld * (V.weightsOnEdge[x * ME2 + j] * AT(u, eoe, k) * workpv);
    }
  }
  FLD(q)[ix] = q;
  const double* ke = FLD(ke); const double* hd = FLD(h_divergence); const double* w = FLD(w);
  tend_u += rho_e * (q - (AT(ke, c2, k) - AT(ke, c1, k)) * invDc) - u_k * 0.5 * (AT(hd, c1, k) + AT(hd, c2, k));   // :1005-1007
  {
    const double wsum = AT(w, c1, k) + AT(w, c1, k + 1) + AT(w, c2, k) + AT(w, c2, k + 1);
    tend_u -= (P.omega2 * V.cosAngleEdge[x] * V.cosLatEdge[x] * rho_e * 0.25 * wsum)
              - (u_k * 0.25 * wsum * rho_e * P.inv_r_earth);                                       // :1011-1017
  }
  double tue;
  if (RK0) {
    const double* pp = FLD(pressure_p); const double* zz = FLD(zz); const double* dpdz = FLD(dpdz);
    tue = -FLD(cqu)[ix] * ((AT(pp, c2, k) - AT(pp, c1, k)) * invDc / (0.5 * (AT(zz, c2, k) + AT(zz, c1, k)))
                           - 0.5 * FLD(zxu)[ix] * (AT(dpdz, c1, k) + AT(dpdz, c2, k)));             // :967-969
    const int v1 = V.verticesOnEdge[x * 2], v2 = V.verticesOnEdge[x * 2 + 1];
    const double* dv = FLD(divergence); const double* vo = FLD(vorticity); const double* kd = FLD(kdiff);
    const double r_dc = invDc;
    const double r_dv = dmin(V.invDvEdge[x], 4 * r_dc);
    const double u_diffusion = (AT(dv, c2, k) - AT(dv, c1, k)) * r_dc - (AT(vo, v2, k) - AT(vo, v1, k)) * r_dv;   // :1041-1042
    const double kdiffu = 0.5 * (AT(kd, c1, k) + AT(kd, c2, k));
    tue += rho_e * kdiffu * u_diffusion * V.meshScalingDel2[x];                                     // :1046-1047
    if (P.visc4_on) {                                                                               // :1072-1090
      const double* dd = FLD(delsq_divergence); const double* dvo = FLD(delsq_vorticity);
      const double u_mix_scale = V.meshScalingDel4[x] * P.h_mom_eddy_visc4;
      const double r_dc4 = u_mix_scale * P.del4u_div_factor * invDc;
      const double r_dv4 = u_mix_scale * dmin(V.invDvEdge[x], 4 * invDc);
      const double ud4 = rho_e * ((AT(dd, c2, k) - AT(dd, c1, k)) * r_dc4 - (AT(dvo, v2, k) - AT(dvo, v1, k)) * r_dv4);
      tue -= ud4;
    }
    if (P.vmix_u_on && k > 0 && k < L - 1) {                                                        // :1094-1146
      const double* zg = FLD(zgrid);
      const double z1 = 0.5 * (AT(zg, c1, k - 1) + AT(zg, c2, k - 1)), z2 = 0.5 * (AT(zg, c1, k) + AT(zg, c2, k));
      const double z3 = 0.5 * (AT(zg, c1, k + 1) + AT(zg, c2, k + 1)), z4 = 0.5 * (AT(zg, c1, k + 2) + AT(zg, c2, k + 2));
      const double zm = 0.5 * (z1 + z2), z0 = 0.5 * (z2 + z3), zp = 0.5 * (z3 + z4);
      double up, uc, um;
      if (P.mix_full) { up = u[ix + 1]; uc = u_k; um = u[ix - 1]; }
      else { up = s_um[k + 1]; uc = s_um[k]; um = s_um[k - 1]; }
      tue += rho_e * P.v_mom_eddy_visc2 * ((up - uc) / (zp - z0) - (uc - um) / (z0 - zm)) / (0.5 * (zp - zm));
    }
    FLD(tend_u_euler)[ix] = tue;
  } else {
    tue = FLD(tend_u_euler)[ix];
  }
  if (P.rayleigh_u && k > L - P.rayleigh_levels + 1) {                                              // :1152-1159
    const double coef = (double)((double)k - (L - P.rayleigh_levels)) * P.rayleigh_coef_inverse;
    tend_u -= rho_e * u_k * coef;
  }
  tend_u += tue + FLD(tend_ru_physics)[ix];                                                         // :1162
  FLD(tend_u)[ix] = tend_u;
}

// helper: horizontal advection + curvature part of the w tendency  :1170-1218  (value of cr.w before mixing)
__device__ __forceinline__ double w_adv_curv(const View& V, const DynTendParams& P, int x, int k, size_t ix, int LP) {
  if (k == 0) return 0.0;
  const int ME = V.maxEdges, NA = V.nAdv, n = V.nEdgesOnCell[x];
  const double* fzm = FLD(fzm); const double* fzp = FLD(fzp);
  const double fm = fzm[k], fp = fzp[k];
  double wv = 0.0;
  if (n > 0) {
    // ru_edge_w / flux_arr are per-point fields overwritten for every edge (Q18): only the LAST
    // edge's values survive, and the second loop multiplies them by every edge's sign.
    const double* ru = FLD(ru);
    const int e = V.edgesOnCell[x * ME + (n - 1)];
    const double rew = fm * AT(ru, e, k) + fp * AT(ru, e, k - 1);
    double fa = 0.0;
    const int na = V.nAdvCellsForEdge[e];
    for (int j = 0; j < na; ++j) {
      const double sw = V.adv_coefs[e * NA + j] + copysign(1.0, rew) * V.adv_coefs_3rd[e * NA + j];
      fa += sw * 0.0;          // cr.w was zeroed on levels < L just before (:1170-1172); the pad cell is zero too
    }
    FLD(ru_edge_w)[ix] = rew;
    for (int i = 0; i < n; ++i) wv -= V.edgesOnCell_sign[x * ME + i] * rew * fa;                   // :1202
  }
  const double* rz = FLD(rho_zz); const double* uz = FLD(uReconstructZonal); const double* um = FLD(uReconstructMeridional);
  const double rzf = rz[ix] * fm + rz[ix - 1] * fp;
  const double uzf = fm * uz[ix] + fp * uz[ix - 1];
  const double umf = fm * um[ix] + fp * um[ix - 1];
  wv += rzf * (uzf * uzf + umf * umf) / P.r_earth + P.omega2 * V.cosLatCell[x] * uzf * rzf;       // :1210-1216
  return wv;
}

// rk_step == 0, cell pass A: w after advection+curvature (:1170-1218) and the first del^2 of theta (:1365-1382)
__global__ void k_dt_cellA(const View V, const DynTendParams P) {
  COLUMN_THREAD(V.nCells)
  if (!inx || k >= L) return;
  FLD(w)[ix] = w_adv_curv(V, P, x, k, ix, LP);
  const int ME = V.maxEdges, n = V.nEdgesOnCell[x];
  const double* tm = FLD(theta_m); const double* kd = FLD(kdiff); const double* re = FLD(rho_edge);
  const double r_areaCell = V.invAreaCell[x];
  double dsq = 0.0, tte = 0.0;
  for (int i = 0; i < n; ++i) {
    const int e = V.edgesOnCell[x * ME + i];
    const double edge_sign = r_areaCell * V.edgesOnCell_sign[x * ME + i] * V.dvEdge[e] * V.invDcEdge[e];
    const double pr_scale = P.prandtl_inv * V.meshScalingDel2[e];
    const int c1 = V.cellsOnEdge[e * 2], c2 = V.cellsOnEdge[e * 2 + 1];
    double flux = edge_sign * (AT(tm, c2, k) - AT(tm, c1, k)) * AT(re, e, k);
    dsq += flux;
    flux *= 0.5 * (AT(kd, c1, k) + AT(kd, c2, k)) * pr_scale;
    tte += flux;
  }
  FLD(delsq_theta)[ix] = dsq;
  FLD(tend_theta_euler)[ix] = tte;
}
// rk_step == 0, cell pass B: first del^2 of w  :1231-1254  (needs pass A's w on neighbour cells)
__global__ void k_dt_cellB(const View V, const DynTendParams P) {
  COLUMN_THREAD(V.nCells)
  if (!inx || k >= L) return;
  double dsq = 0.0, twe = 0.0;
  if (k > 0) {
    const int ME = V.maxEdges, n = V.nEdgesOnCell[x];
    const double* w = FLD(w); const double* kd = FLD(kdiff); const double* re = FLD(rho_edge);
    const double r_areaCell = V.invAreaCell[x];
    for (int i = 0; i < n; ++i) {
      const int e = V.edgesOnCell[x * ME + i];
      const double edge_sign = 0.5 * r_areaCell * V.edgesOnCell_sign[x * ME + i] * V.dvEdge[e] * V.invDcEdge[e];
      const int c1 = V.cellsOnEdge[e * 2], c2 = V.cellsOnEdge[e * 2 + 1];
      double flux = edge_sign * (AT(re, e, k) + AT(re, e, k - 1)) * (AT(w, c2, k) - AT(w, c1, k));
      dsq += flux;
      flux *= V.meshScalingDel2[e] * 0.25 * (AT(kd, c1, k) + AT(kd, c2, k) + AT(kd, c1, k - 1) + AT(kd, c2, k - 1));
      twe += flux;
    }
  }
  FLD(delsq_w)[ix] = dsq;
  FLD(tend_w_euler)[ix] = twe;
}
// final cell pass: w (:1256-1322) and theta (:1328-1479)
template <bool RK0>
__global__ void k_dt_cellC(const View V, const DynTendParams P) {
  extern __shared__ double sm[];
  COLUMN_THREAD(V.nCells)
  const int TS = LP + 1;
  double* s_a = sm + (size_t)threadIdx.y * TS;                       // w, later wdtz
  double* s_b = sm + (size_t)(blockDim.y + threadIdx.y) * TS;        // wdwz, later post-multiply w
  const bool act = inx && k < L;
  const int ME = V.maxEdges, NA = V.nAdv;
  const int n = inx ? V.nEdgesOnCell[x] : 0;
  const double* fzm = FLD(fzm); const double* fzp = FLD(fzp); const double* rdzu = FLD(rdzu); const double* rdzw = FLD(rdzw);
  const double* rw = FLD(rw);
  double w_k = 0.0, twe = 0.0;
  if (act) {
    if (RK0) {
      w_k = FLD(w)[ix];
      twe = FLD(tend_w_euler)[ix];
      if (P.visc4_on && k > 0) {                                                                     // :1256-1272
        const double* dsw = FLD(delsq_w);
        const double r_areaCell = P.h_mom_eddy_visc4 * V.invAreaCell[x];
        for (int i = 0; i < n; ++i) {
          const int e = V.edgesOnCell[x * ME + i];
          const int c1 = V.cellsOnEdge[e * 2], c2 = V.cellsOnEdge[e * 2 + 1];
          const double edge_sign = V.meshScalingDel4[e] * r_areaCell * V.dvEdge[e] * V.edgesOnCell_sign[x * ME + i] * V.invDcEdge[e];
          twe -= edge_sign * (AT(dsw, c2, k) - AT(dsw, c1, k));
        }
      }
    } else {
      w_k = w_adv_curv(V, P, x, k, ix, LP);
    }
    s_a[k] = w_k;
  }
  if (inx && k == L) { s_a[L] = FLD(w)[ix]; s_b[L] = FLD(wdwz)[ix]; }   // level L keeps its stored value
  __syncthreads();
  double wdwz_k = 0.0;
  if (act) {                                                                                          // :1277-1287
    if (k == 1 || k == L - 1) wdwz_k = 0.25 * (rw[ix] + rw[ix - 1]) * (s_a[k] + s_a[k - 1]);
    if (k > 1 && k < L - 1) wdwz_k = flux3(s_a[k - 2], s_a[k - 1], s_a[k], s_a[k + 1], 0.5 * (rw[ix] + rw[ix - 1]), 1.0);
    s_b[k] = wdwz_k;
    FLD(wdwz)[ix] = wdwz_k;
  }
  __syncthreads();
  double wdwz_p = 0.0;
  if (act) wdwz_p = s_b[k + 1];
  __syncthreads();
  if (act) {
    if (k > 0) w_k *= V.invAreaCell[x] - rdzu[k] * (wdwz_p - wdwz_k);                                // :1292
    if (RK0 && k > 0) {                                                                               // :1297-1299
      const double* pp = FLD(pressure_p); const double* dpdz = FLD(dpdz);
      twe -= FLD(cqw)[ix] * (rdzu[k] * (pp[ix] - pp[ix - 1]) - (fzm[k] * dpdz[ix] + fzp[k] * dpdz[ix - 1]));
    }
    s_b[k] = w_k;                                                    // post-multiply w, for the vertical mixing of w
  }
  if (inx && k == L) s_b[L] = s_a[L];
  __syncthreads();
  if (act) {
    if (RK0 && P.vmix_u_on && k > 0) {                                                                // :1304-1314
      const double* rz = FLD(rho_zz);
      twe += P.v_mom_eddy_visc2 * (rz[ix] + rz[ix - 1]) * 0.5
             * ((s_b[k + 1] - s_b[k]) * rdzw[k] - (s_b[k] - s_b[k - 1]) * rdzw[k - 1]) * rdzu[k];
    }
    if (!RK0 && k > 0) twe = FLD(tend_w_euler)[ix];
    if (RK0) FLD(tend_w_euler)[ix] = twe;
    if (k > 0) w_k += twe;                                                                            // :1320
    FLD(w)[ix] = w_k;
  }
  // ---------------- theta ----------------
  const double* tm = FLD(theta_m); const double* tms = FLD(theta_m_save); const double* rws = FLD(rw_save);
  double tt = 0.0, wdtz_k = 0.0;
  if (act) {
    const double* ru = FLD(ru);
    double fa_last = 0.0;
    for (int i = 0; i < n; ++i) {                                                                     // :1328-1344
      const int e = V.edgesOnCell[x * ME + i];
      const double ru_e = AT(ru, e, k);
      const int na = V.nAdvCellsForEdge[e];
      double fa = 0.0;
      for (int j = 0; j < na; ++j) {
        const int ac = V.advCellsForEdge[e * NA + j];
        const double sw = V.adv_coefs[e * NA + j] + copysign(1.0, ru_e) * V.adv_coefs_3rd[e * NA + j];
        fa += sw * AT(tm, ac, k);
      }
      tt -= V.edgesOnCell_sign[x * ME + i] * ru_e * fa;
      fa_last = fa;
    }
    if (n > 0) FLD(flux_arr)[ix] = fa_last;
    if (P.rk_step > 0) {                                                                              // :1347-1360
      const double* rus = FLD(ru_save);
      for (int i = 0; i < n; ++i) {
        const int e = V.edgesOnCell[x * ME + i];
        const int c1 = V.cellsOnEdge[e * 2], c2 = V.cellsOnEdge[e * 2 + 1];
        const double flux = V.edgesOnCell_sign[x * ME + i] * V.dvEdge[e] * (AT(rus, e, k) - AT(ru, e, k)) * 0.5
                            * (AT(tms, c2, k) + AT(tms, c1, k));
        tt -= flux;
      }
    }
    if (k > 0 && k < L - 1) wdtz_k = ((rws[ix] - rw[ix]) * (fzm[k] * tms[ix] + fzp[k] * tms[ix - 1]));   // :1409-1412
    if (k == 1) wdtz_k += rw[ix] * (fzm[k] * tm[ix] + fzp[k] * tm[ix - 1]);                           // :1413-1415
    if (k == L - 1) wdtz_k = rws[ix] * (fzm[k] * tms[ix] + fzp[k] * tms[ix - 1]);                     // :1416-1419
    s_a[k] = wdtz_k;
    FLD(wdtz)[ix] = wdtz_k;
  }
  if (inx && k == L) s_a[L] = FLD(wdtz)[ix];
  __syncthreads();
  if (!act) return;
  const double rz = FLD(rho_zz)[ix];
  tt *= V.invAreaCell[x] - rdzw[k] * (s_a[k + 1] - wdtz_k);                                          // :1423
  FLD(tend_rtheta_adv)[ix] = tt;
  FLD(rthdynten)[ix] = tt / rz;
  tt += rz * FLD(rt_diabatic_tend)[ix];
  double tte = FLD(tend_theta_euler)[ix];
  if (RK0) {
    if (P.visc4_on) {                                                                                 // :1384-1399
      const double* dst = FLD(delsq_theta);
      const double r_areaCell = P.h_theta_eddy_visc4 * P.prandtl_inv * V.invAreaCell[x];
      for (int i = 0; i < n; ++i) {
        const int e = V.edgesOnCell[x * ME + i];
        const double edge_sign = V.meshScalingDel4[e] * r_areaCell * V.dvEdge[e] * V.edgesOnCell_sign[x * ME + i] * V.invDcEdge[e];
        const int c1 = V.cellsOnEdge[e * 2], c2 = V.cellsOnEdge[e * 2 + 1];
        tte -= edge_sign * (AT(dst, c2, k) - AT(dst, c1, k));
      }
    }
    if (P.vmix_t_on && k > 0 && k < L - 1) {                                                          // :1432-1473
      const double* zg = FLD(zgrid);
      const double z1 = zg[ix - 1], z2 = zg[ix], z3 = zg[ix + 1], z4 = zg[ix + 2];
      const double zm = 0.5 * (z1 + z2), z0 = 0.5 * (z2 + z3), zp = 0.5 * (z3 + z4);
      if (P.mix_full) {
        tte += P.v_theta_eddy_visc2 * P.prandtl_inv * rz * ((tm[ix + 1] - tm[ix]) / (zp - z0) - (tm[ix] - tm[ix - 1]) / (z0 - zm)) / (0.5 * (zp - zm));
      } else {
        const double* ti = FLD(t_init);
        tte += P.v_theta_eddy_visc2 * P.prandtl_inv * rz
               * (((tm[ix + 1] - ti[ix + 1]) - (tm[ix] - ti[ix])) / (zp - z0) - ((tm[ix] - ti[ix]) - (tm[ix - 1] - ti[ix - 1])) / (z0 - zm)) / (0.5 * (zp - zm));
      }
    }
    FLD(tend_theta_euler)[ix] = tte;
  }
  tt += tte + FLD(tend_rtheta_physics)[ix];                                                           // :1478
  FLD(tend_theta)[ix] = tt;
}

// ============================================================================================
// atm_set_smlstep_pert_variables_work  :1503-1528  (levels 0..L-1 of the cells of cpr; level -1 reads 0)
__global__ void k_smlstep(const View V, int nRelaxZone) {
  COLUMN_THREAD(V.nCells)
  if (!inx || k >= L) return;
  if (!V.inCpr[x] || V.bdyMaskCell[x] > nRelaxZone) return;
  const int ME = V.maxEdges, n = V.nEdgesOnCell[x];
  const double* ut = FLD(u_tend); const double* zb = FLD(zb_cell); const double* zb3 = FLD(zb3_cell);
  const double fm = FLD(fzm)[k], fp = FLD(fzp)[k];
  double wv = FLD(w)[ix];
  for (int i = 0; i < n; ++i) {
    const int e = V.edgesOnCell[x * ME + i];
    const double ut_k = AT(ut, e, k);
    const double ut_m = (k > 0) ? AT(ut, e, k - 1) : 0.0;
    const double flux = V.edgesOnCell_sign[x * ME + i] * (fm * ut_k + fp * ut_m);
    wv -= (zb[i * V.cellSlot + ix] + copysign(1.0, ut_k) * zb3[i * V.cellSlot + ix]) * flux;
  }
  const double zz_k = FLD(zz)[ix];
  const double zz_m = (k > 0) ? FLD(zz)[ix - 1] : 0.0;
  wv *= (fm * zz_k + fp * zz_m);
  FLD(w)[ix] = wv;
}

// ============================================================================================
// atm_advance_acoustic_step_work  :1546-1705
// phase 1 (all levels in parallel): rtheta_pp_old (:1615-1623), zeroing of level L (:1625-1630),
// horizontal flux parts of rs/ts (:1644-1652) into scratch.
__global__ void k_acoustic_flux(const View V, double dts, int small_step) {
  COLUMN_THREAD(V.nCells)
  if (!inx) return;
  if (k == L && small_step == 0) { FLD(wwAvg)[ix] = 0; FLD(rw_p)[ix] = 0; }
  if (k >= L) return;
  FLD(rtheta_pp_old)[ix] = (small_step == 0) ? 0.0 : FLD(rtheta_pp)[ix];
  if (V.specZoneMaskCell[x] != 0.0) return;
  const int ME = V.maxEdges, n = V.nEdgesOnCell[x];
  const double* ru_p = FLD(ru_p); const double* tm = FLD(theta_m);
  const double inva = V.invAreaCell[x];
  double rs = 0, ts = 0;
  for (int i = 0; i < n; ++i) {
    const int e = V.edgesOnCell[x * ME + i];
    const int c1 = V.cellsOnEdge[e * 2], c2 = V.cellsOnEdge[e * 2 + 1];
    const double flux = V.edgesOnCellSign[x * ME + i] * dts * V.dvEdge[e] * AT(ru_p, e, k) * inva;
    rs -= flux;
    ts -= flux * 0.5 * (AT(tm, c2, k) + AT(tm, c1, k));
  }
  V.scr_rs[ix] = rs; V.scr_ts[ix] = ts;
}
// phase 2 (one thread per column, levels ascending): the vertically implicit sweep  :1657-1703.
// rs[k-1] and ts[k-1] are always 0 (the reference re-zeroes both arrays at every point, Q25); the
// back-substitution is absent (Q28); cr.theta_m stands in for tend_rt and cr.w for tend_rw (Q27).
__global__ void k_acoustic_column(const View V, double dts, int small_step, double epssm, double resm) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= V.nCells) return;
  const int L = V.L, LP = V.LP;
  const size_t b = (size_t)c * LP;
  const bool S0 = small_step == 0;
  const double* cofrz = FLD(cofrz); const double* rdzw = FLD(rdzw); const double* fzm = FLD(fzm); const double* fzp = FLD(fzp);
  double* rho_pp = FLD(rho_pp) + b; double* rtheta_pp = FLD(rtheta_pp) + b; double* rw_p = FLD(rw_p) + b; double* wwAvg = FLD(wwAvg) + b;
  const double* tend_rho = FLD(tend_rho) + b; const double* theta_m = FLD(theta_m) + b; const double* w = FLD(w) + b;
  const double* coftz = FLD(coftz) + b; const double* cofwz = FLD(cofwz) + b; const double* cofwr = FLD(cofwr) + b; const double* cofwt = FLD(cofwt) + b;
  const double* a_tri = FLD(a_tri) + b; const double* alpha_tri = FLD(alpha_tri) + b; const double* zz = FLD(zz) + b;
  const double* rw_save = FLD(rw_save) + b; const double* rw = FLD(rw) + b; const double* dss = FLD(dss) + b; const double* rho_zz = FLD(rho_zz) + b;
  const double* srs = V.scr_rs + b; const double* sts = V.scr_ts + b;
  const bool spec = V.specZoneMaskCell[c] != 0.0;
  double rw_prev = 0, rho_prev = 0, rt_prev = 0;
  double rw_old_k = S0 ? 0.0 : rw_p[0];
  double zz_m = 0.0, cofwt_m = 0.0, rz_m = 0.0;
  for (int k = 0; k < L; ++k) {
    const double rw_old_p = S0 ? 0.0 : rw_p[k + 1];
    const double rho_old = S0 ? 0.0 : rho_pp[k];
    const double rt_old = S0 ? 0.0 : rtheta_pp[k];
    const double ww_old = S0 ? 0.0 : wwAvg[k];
    const double zz_k = zz[k], cofwt_k = cofwt[k], rz_k = rho_zz[k];
    double rw_new, rho_new, rt_new, ww_new = ww_old;
    if (!spec) {
      const double coftz_k = coftz[k], coftz_p = coftz[k + 1];
      const double rs = rho_old + dts * tend_rho[k] + srs[k] - cofrz[k] * resm * (rw_old_p - rw_old_k);                       // :1657
      const double ts = rt_old + dts * theta_m[k] + sts[k] - resm * rdzw[k] * (coftz_p * rw_old_p - coftz_k * rw_old_k);       // :1658
      rw_new = rw_old_k;
      if (k > 0) {
        const double w_k = w[k];
        ww_new += 0.5 * (1.0 - epssm) * rw_old_k;                                                                            // :1661
        rw_new += dts * w_k - cofwz[k] * ((zz_k * ts - zz_m * 0.0) + resm * (zz_k * rt_old - zz_m * rt_prev))
                  - cofwr[k] * ((rs + 0.0) + resm * (rho_old + rho_prev))
                  + cofwt_k * (ts + resm * rt_old)
                  + cofwt_m * (0.0 + resm * rt_prev);                                                                         // :1662-1667
        rw_new -= a_tri[k] * rw_prev;                                                                                        // :1670
        rw_new *= alpha_tri[k];                                                                                              // :1671
        const double dsk = dss[k];
        const double r3 = rw_save[k] - rw[k];
        rw_new += r3 - dts * dsk * (fzm[k] * zz_k + fzp[k] * zz_m) * (fzm[k] * rz_k + fzp[k] * rz_m) * w_k;                   // :1682-1684
        rw_new /= (1.0 + dts * dsk);                                                                                         // :1685
        rw_new -= r3;                                                                                                        // :1686
        ww_new += 0.5 * (1.0 + epssm) * rw_new;                                                                              // :1689
      }
      rho_new = rs - cofrz[k] * (rw_old_p - rw_new);                                                                         // :1694
      rt_new = ts - rdzw[k] * (coftz_p * rw_old_p - coftz_k * rw_new);                                                       // :1695-1696
    } else {                                                                                                                 // :1698-1703
      rho_new = rho_old + dts * tend_rho[k];
      rt_new = rt_old + dts * theta_m[k];
      rw_new = rw_old_k + dts * w[k];
      ww_new = ww_old + 0.5 * (1.0 + epssm) * rw_new;
    }
    rho_pp[k] = rho_new; rtheta_pp[k] = rt_new;
    if (S0 || spec || k > 0) { rw_p[k] = rw_new; wwAvg[k] = ww_new; }
    rw_prev = rw_new; rho_prev = rho_new; rt_prev = rt_new;
    rw_old_k = rw_old_p; zz_m = zz_k; cofwt_m = cofwt_k; rz_m = rz_k;
  }
}

// Fused single-kernel form of the acoustic step (default).  lane = level, a block owns whole columns.
// Everything except the dependence of rw_p(k) on the freshly updated level k-1 is evaluated in
// parallel exactly as written in the reference.  That dependence is affine: with x = rw_p_new,
//     rho_pp_new(k-1)    = rp0 + rp1 * x(k-1)          (:1694)
//     rtheta_pp_new(k-1) = rt0 + rt1 * x(k-1)          (:1695-1696)
//     x(k) = P(k) + Q(k) * x(k-1)                      (:1662-1686 collected in x(k-1))
// so every thread computes its (P, Q) and ONE thread per column runs the 1-multiply-add-per-level
// sweep out of shared memory; rho_pp / rtheta_pp / wwAvg then follow in parallel from x with the
// reference's own expressions.  Differs from the literal left-to-right evaluation only by the
// regrouping of terms inside one level (a few ulp; checked against the oracle at 1e-12).
template <bool S0>
__global__ void k_acoustic(const View V, double dts, double epssm, double resm) {
  extern __shared__ double sm[];
  COLUMN_THREAD(V.nCells)
  const int TS = LP + 1;
  double* s_rp0 = sm + (size_t)threadIdx.y * TS;
  double* s_rt0 = sm + (size_t)(blockDim.y + threadIdx.y) * TS;
  double* s_P = sm + (size_t)(2 * blockDim.y + threadIdx.y) * TS;
  double* s_Q = sm + (size_t)(3 * blockDim.y + threadIdx.y) * TS;
  const bool act = inx && k < L;
  const bool spec = inx ? (V.specZoneMaskCell[x] != 0.0) : false;
  if (inx && k == L && S0) { FLD(wwAvg)[ix] = 0; FLD(rw_p)[ix] = 0; }                                // :1625-1630, level L
  double rs = 0, ts = 0, rw_old_k = 0, rw_old_p = 0, rho_old = 0, rt_old = 0, ww_old = 0;
  double coftz_k = 0, coftz_p = 0, cofrz_k = 0, rdzw_k = 0, w_k = 0, tr_k = 0, tm_k = 0;
  if (act) {
    const double* tm = FLD(theta_m);
    cofrz_k = FLD(cofrz)[k]; rdzw_k = FLD(rdzw)[k];
    if (!S0) {
      rw_old_k = FLD(rw_p)[ix]; rw_old_p = FLD(rw_p)[ix + 1];
      rho_old = FLD(rho_pp)[ix]; rt_old = FLD(rtheta_pp)[ix]; ww_old = FLD(wwAvg)[ix];
    }
    FLD(rtheta_pp_old)[ix] = S0 ? 0.0 : rt_old;                                                       // :1615-1623
    w_k = FLD(w)[ix]; tr_k = FLD(tend_rho)[ix]; tm_k = tm[ix];
    if (!spec) {
      const int ME = V.maxEdges, n = V.nEdgesOnCell[x];
      const double* ru_p = FLD(ru_p);
      const double inva = V.invAreaCell[x];
      for (int i = 0; i < n; ++i) {                                                                   // :1644-1652
        const int e = V.edgesOnCell[x * ME + i];
        const int c1 = V.cellsOnEdge[e * 2], c2 = V.cellsOnEdge[e * 2 + 1];
        const double flux = V.edgesOnCellSign[x * ME + i] * dts * V.dvEdge[e] * AT(ru_p, e, k) * inva;
        rs -= flux;
        ts -= flux * 0.5 * (AT(tm, c2, k) + AT(tm, c1, k));
      }
      coftz_k = FLD(coftz)[ix]; coftz_p = FLD(coftz)[ix + 1];
      rs = rho_old + dts * tr_k + rs - cofrz_k * resm * (rw_old_p - rw_old_k);                        // :1657
      ts = rt_old + dts * tm_k + ts - resm * rdzw_k * (coftz_p * rw_old_p - coftz_k * rw_old_k);      // :1658
      // new rho_pp / rtheta_pp of THIS level as affine functions of x(k):  rp0 + cofrz*x,  rt0 + rdzw*coftz*x
      s_rp0[k] = rs - cofrz_k * rw_old_p;
      s_rt0[k] = ts - rdzw_k * (coftz_p * rw_old_p);
    }
  }
  __syncthreads();
  double r3 = 0, r2 = 1, cofwt_k = 0, zz_k = 0;
  if (act && !spec) {
    zz_k = FLD(zz)[ix]; cofwt_k = FLD(cofwt)[ix];
    if (k == 0) { s_P[0] = rw_old_k; s_Q[0] = 0.0; }
    else {
      const double zz_m = FLD(zz)[ix - 1], cofwt_m = FLD(cofwt)[ix - 1], rz_k = FLD(rho_zz)[ix], rz_m = FLD(rho_zz)[ix - 1];
      const double cofwz_k = FLD(cofwz)[ix], cofwr_k = FLD(cofwr)[ix];
      const double fm = FLD(fzm)[k], fp = FLD(fzp)[k];
      const double dsk = FLD(dss)[ix];
      r3 = FLD(rw_save)[ix] - FLD(rw)[ix];
      const double r1 = r3 - dts * dsk * (fm * zz_k + fp * zz_m) * (fm * rz_k + fp * rz_m) * w_k;      // :1682-1684
      r2 = 1.0 + dts * dsk;                                                                           // :1685
      // terms of :1662-1667 that do not involve level k-1's new values
      const double A0 = rw_old_k + (dts * w_k - cofwz_k * ((zz_k * ts - zz_m * 0.0) + resm * (zz_k * rt_old))
                                    - cofwr_k * ((rs + 0.0) + resm * rho_old) + cofwt_k * (ts + resm * rt_old));
      const double A1 = resm * (cofwz_k * zz_m + cofwt_m);        // coefficient of rtheta_pp_new(k-1)
      const double A2 = resm * cofwr_k;                           // coefficient of -rho_pp_new(k-1)
      const double rp0 = s_rp0[k - 1], rt0 = s_rt0[k - 1];
      const double rp1 = FLD(cofrz)[k - 1], rt1 = FLD(rdzw)[k - 1] * FLD(coftz)[ix - 1];
      const double B0 = A0 + A1 * rt0 - A2 * rp0;
      const double B1 = A1 * rt1 - A2 * rp1 - FLD(a_tri)[ix];                                          // :1670
      const double al = FLD(alpha_tri)[ix];                                                            // :1671
      s_P[k] = (B0 * al + r1) / r2 - r3;                                                               // :1682-1686
      s_Q[k] = B1 * al / r2;
    }
  }
  __syncthreads();
  if (inx && !spec && k == 0) {             // the sweep: one multiply-add per level, levels ascending (M4)
    double xv = s_P[0];
    for (int kk = 1; kk < L; ++kk) { xv = s_P[kk] + s_Q[kk] * xv; s_P[kk] = xv; }
  }
  __syncthreads();
  if (!act) return;
  double rw_new, rho_new, rt_new, ww_new = ww_old;
  if (!spec) {
    rw_new = s_P[k];
    if (k > 0) {
      ww_new += 0.5 * (1.0 - epssm) * rw_old_k;                                                        // :1661
      ww_new += 0.5 * (1.0 + epssm) * rw_new;                                                          // :1689
    }
    rho_new = rs - cofrz_k * (rw_old_p - rw_new);                                                      // :1694
    rt_new = ts - rdzw_k * (coftz_p * rw_old_p - coftz_k * rw_new);                                    // :1695-1696
  } else {                                                                                             // :1698-1703
    rho_new = rho_old + dts * tr_k;
    rt_new = rt_old + dts * tm_k;
    rw_new = rw_old_k + dts * w_k;
    ww_new = ww_old + 0.5 * (1.0 + epssm) * rw_new;
  }
  FLD(rho_pp)[ix] = rho_new; FLD(rtheta_pp)[ix] = rt_new;
  if (S0 || spec || k > 0) { FLD(rw_p)[ix] = rw_new; FLD(wwAvg)[ix] = ww_new; }
}

// ============================================================================================
// atm_divergence_damping_3d  :1726-1763
__global__ void k_divdamp(const View V, double coef_divdamp) {
  COLUMN_THREAD(V.nEdges)
  if (!inx || k >= L) return;
  const int c1 = V.cellsOnEdge[x * 2], c2 = V.cellsOnEdge[x * 2 + 1];
  if (V.isShared[c1] && V.isShared[c2]) return;
  const double* rpp = FLD(rtheta_pp); const double* rppo = FLD(rtheta_pp_old); const double* tm = FLD(theta_m);
  const double divCell1 = -(AT(rpp, c1, k) - AT(rppo, c1, k));
  const double divCell2 = -(AT(rpp, c2, k) - AT(rppo, c2, k));
  FLD(ru_p)[ix] += coef_divdamp * (divCell2 - divCell1) * (1.0 - V.specZoneMaskEdge[x]) / (AT(tm, c1, k) + AT(tm, c2, k));
}

// ============================================================================================
// atm_recover_large_step_variables_work  :1766-1872
__global__ void k_rec_pad(const View V) {           // :1792-1794: rho_zz = 1 on the "garbage cell" = the pad cell
  const int k = threadIdx.x;
  if (k < V.L) FLD(rho_zz)[(size_t)V.nCells * V.LP + k] = 1.0;
}
__global__ void k_rec_cell1(const View V, double invNs, int rk_step, double dt, double rgas, double rcv) {   // :1800-1826
  COLUMN_THREAD(V.nCells)
  if (!inx || k >= L) return;
  const double rho_p = FLD(rho_p_save)[ix] + FLD(rho_pp)[ix];
  const double rho_zz = rho_p + FLD(rho_base)[ix];
  FLD(rho_p)[ix] = rho_p; FLD(rho_zz)[ix] = rho_zz;
  double ww = FLD(wwAvg)[ix];
  ww *= invNs; ww += FLD(rw_save)[ix];
  FLD(wwAvg)[ix] = ww;
  const double rwv = FLD(rw_save)[ix] + FLD(rw_p)[ix];
  FLD(rw)[ix] = rwv;
  const double zz_k = FLD(zz)[ix];
  const double zz_m = (k > 0) ? FLD(zz)[ix - 1] : 0.0;
  FLD(w)[ix] = rwv / (FLD(fzm)[k] * zz_k + FLD(fzp)[k] * zz_m);                                      // :1810
  const double rtb = FLD(rtheta_base)[ix];
  if (rk_step == 2) {
    const double rtp = FLD(rtheta_p_save)[ix] + FLD(rtheta_pp)[ix] - dt * rho_zz * FLD(rt_diabatic_tend)[ix];
    FLD(rtheta_p)[ix] = rtp;
    FLD(theta_m)[ix] = (rtp + rtb) / rho_zz;
    const double ex = zz_k * (rgas / 100000) * pow((rtp + rtb), rcv);                                 // :1819
    FLD(exner)[ix] = ex;
    FLD(pressure_p)[ix] = zz_k * rgas * (ex * rtp + rtb * (ex - FLD(exner_base)[ix]));               // :1821
  } else {
    const double rtp = FLD(rtheta_p_save)[ix] + FLD(rtheta_pp)[ix];
    FLD(rtheta_p)[ix] = rtp;
    FLD(theta_m)[ix] = (rtp + rtb) / rho_zz;
  }
}
__global__ void k_rec_edge(const View V, double invNs) {                                               // :1835-1842
  COLUMN_THREAD(V.nEdges)
  if (!inx || k >= L) return;
  const int c1 = V.cellsOnEdge[x * 2], c2 = V.cellsOnEdge[x * 2 + 1];
  const double* rz = FLD(rho_zz);
  double ra = FLD(ruAvg)[ix];
  ra *= invNs; ra += FLD(ru_save)[ix];
  FLD(ruAvg)[ix] = ra;
  const double ruv = FLD(ru_save)[ix] * FLD(ru_p)[ix];                                                // a product, as written (:1840)
  FLD(ru)[ix] = ruv;
  FLD(u)[ix] = 2 * ruv / (AT(rz, c1, k) + AT(rz, c2, k));
}
__global__ void k_rec_cell2(const View V, int nRelaxZone) {                                            // :1844-1871
  COLUMN_THREAD(V.nCells)
  if (!inx || k >= L) return;
  if (V.bdyMaskCell[x] > nRelaxZone) return;
  const int ME = V.maxEdges, n = V.nEdgesOnCell[x];
  const double* ru = FLD(ru); const double* zb = FLD(zb_cell); const double* zb3 = FLD(zb3_cell); const double* rz = FLD(rho_zz);
  const double fm = FLD(fzm)[k], fp = FLD(fzp)[k];
  const double cf1 = FLD(cf1)[0], cf2 = FLD(cf2)[0], cf3 = FLD(cf3)[0];
  double wv = FLD(w)[ix];
  if (k == 0) {
    // the surface term is accumulated once per (cell, LEVEL) iteration: L times, interleaved with
    // the level-0 flux2 term on the first pass (level -1 reads 0)
    for (int kk = 0; kk < L; ++kk) {
      for (int i = 0; i < n; ++i) {
        const int e = V.edgesOnCell[x * ME + i];
        const double sgn = V.edgesOnCell_sign[x * ME + i];
        const double flux = (cf1 * AT(ru, e, 0) + cf2 * AT(ru, e, 1) + cf3 * AT(ru, e, 2));
        wv += sgn * (zb[i * V.cellSlot + ix] + copysign(1.0, flux) * zb3[i * V.cellSlot + ix]) * flux;
        if (kk == 0) {
          const double flux2 = fm * AT(ru, e, 0) * (fp * 0.0);
          wv += sgn * (zb[i * V.cellSlot + ix] + copysign(1.0, flux2) * zb3[i * V.cellSlot + ix]) * flux2;
        }
      }
    }
    wv /= (cf1 * rz[ix] + cf2 * rz[ix + 1] + cf3 * rz[ix + 2]);
  } else {
    for (int i = 0; i < n; ++i) {
      const int e = V.edgesOnCell[x * ME + i];
      const double flux2 = fm * AT(ru, e, k) * (fp * AT(ru, e, k - 1));                                // a product, as written (:1855)
      wv += V.edgesOnCell_sign[x * ME + i] * (zb[i * V.cellSlot + ix] + copysign(1.0, flux2) * zb3[i * V.cellSlot + ix]) * flux2;
    }
    wv /= (fm * rz[ix] + fp * rz[ix - 1]);
  }
  FLD(w)[ix] = wv;
}

// ============================================================================================
// atm_rk_dynamics_substep_finish  :1951-2007
__global__ void k_finish_cell(const View V, int lt_split, int first, int last, double inv_split) {
  COLUMN_THREAD(V.nCells)
  if (!inx || k >= L) return;
  if (lt_split) {
    FLD(rw_save)[ix] = FLD(rw)[ix]; FLD(rtheta_p_save)[ix] = FLD(rtheta_p)[ix]; FLD(rho_p_save)[ix] = FLD(rho_p)[ix];
    FLD(w)[ix] = FLD(w_2)[ix]; FLD(theta_m)[ix] = FLD(theta_m_2)[ix]; FLD(rho_zz)[ix] = FLD(rho_zz_2)[ix];
  }
  double ws = first ? FLD(wwAvg)[ix] : FLD(wwAvg)[ix] + FLD(wwAvg_split)[ix];
  FLD(wwAvg_split)[ix] = ws;
  if (last) { FLD(wwAvg)[ix] = ws * inv_split; FLD(rho_zz)[ix] = FLD(rho_zz_old_split)[ix]; }
}
__global__ void k_finish_edge(const View V, int lt_split, int first, int last, double inv_split) {
  COLUMN_THREAD(V.nEdges)
  if (!inx || k >= L) return;
  if (lt_split) { FLD(ru_save)[ix] = FLD(ru)[ix]; FLD(u)[ix] = FLD(u_2)[ix]; }
  double rs = first ? FLD(ruAvg)[ix] : FLD(ruAvg)[ix] + FLD(ruAvg_split)[ix];
  FLD(ruAvg_split)[ix] = rs;
  if (last) FLD(ruAvg)[ix] = rs * inv_split;
}

// ============================================================================================
// region <-> mirror transfers and halo pack/unpack.  `map[i]` = internal (SFC) index of caller index i.
// staging layout: [i][L1][slots] (exactly the host array of an array-typed region field).
__global__ void k_stage_to_field(double* __restrict__ field, const double* __restrict__ staging, const int* __restrict__ map,
                                 int n, int L1, int LP, int slots, size_t slotStride) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)n * L1 * slots;
  if (t >= total) return;
  const int s = (int)(t % slots); const size_t r = t / slots; const int k = (int)(r % L1); const int i = (int)(r / L1);
  field[s * slotStride + (size_t)map[i] * LP + k] = staging[t];
}
__global__ void k_field_to_stage(const double* __restrict__ field, double* __restrict__ staging, const int* __restrict__ map,
                                 int n, int L1, int LP, int slots, size_t slotStride) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)n * L1 * slots;
  if (t >= total) return;
  const int s = (int)(t % slots); const size_t r = t / slots; const int k = (int)(r % L1); const int i = (int)(r / L1);
  staging[t] = field[s * slotStride + (size_t)map[i] * LP + k];
}
// halo buffers: [field][i][L1]; idx holds internal indices
struct PackArgs { double* f[32]; int nf; };
__global__ void k_pack(const PackArgs A, const int* __restrict__ idx, int n, int L1, int LP, double* __restrict__ buf) {
  const int k = threadIdx.x; const int i = blockIdx.x * blockDim.y + threadIdx.y; const int fi = blockIdx.y;
  if (i >= n || k >= L1) return;
  buf[((size_t)fi * n + i) * L1 + k] = A.f[fi][(size_t)idx[i] * LP + k];
}
__global__ void k_unpack(const PackArgs A, const int* __restrict__ idx, int n, int L1, int LP, const double* __restrict__ buf) {
  const int k = threadIdx.x; const int i = blockIdx.x * blockDim.y + threadIdx.y; const int fi = blockIdx.y;
  if (i >= n || k >= L1) return;
  A.f[fi][(size_t)idx[i] * LP + k] = buf[((size_t)fi * n + i) * L1 + k];
}
