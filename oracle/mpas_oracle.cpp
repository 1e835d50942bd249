// mpas_oracle.cpp -- CPU restatement of the RK3 dynamics hot path of alexaiken/mpas-regent.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.  The product path
// (libmpas_b200.so) never links, calls or falls back to anything in this directory.
//
// PARITY IS UNPINNED BY THE REFERENCE: the reference ships no tests, golden vectors or
// known-answer outputs for this path, and its toolchain (Regent/Terra/
// Legion/libnetcdf) is absent here, so it cannot be built or run (oracle/_ref does not
// exist).  What pins this file instead: (i) it is a literal, loop-by-loop restatement of the
// task bodies, each function citing the lines it follows; (ii) the survey-derived
// cross-checks on the bundled x1.2562 mesh (partition set sizes, pad-index counts) and the
// only observable in output.txt ("Horizontal normal velocity at Edge 0 is 0.000000") are
// reproduced in tests/test_oracle.py; (iii) tests/test_oracle_restatements.py re-derives every non-trivial task from the
// reference text as array-at-a-time numpy, independently of this file's loop form, and requires agreement to 1e-13.
//
// Memory model (the reference leaves it undefined):
//   M1  every field byte is 0 until written.
//   M2  connectivity arrives RESOLVED: 0-based indices with a zero pad entity at index N
//       (LITERAL = stored 1-based id used as index, id==N -> pad; CORRECTED = id-1, id 0 -> pad).
//       Deviation from a literal Legion run: (N,k) would alias into the next level under
//       Legion's default layout; here it reads the pad.  The task bodies are policy-agnostic.
//   M3  level -1 (dynamics_tasks.rg:1520,1525,1810,1855) is a zero pad row.
//   M4  levels ascend within a column (the only order any hot-path loop depends on:
//       rw_p[k-1], rtheta_pp[k-1], rho_pp[k-1] at dynamics_tasks.rg:1663-1670).
//   `cpr` (private_1[i]) is a set of CELLS; atm_set_smlstep_pert_variables runs on EVERY level
//       0..nVertLevels of them (`for iCell in cpr`, dynamics_tasks.rg:1516); atm_divergence_damping_3d indexes it with any cell.
//   Each 3-D field is stored [entity 0..N][level -1..nVertLevels][slot].
//
// Floating point: build with -O2 -ffp-contract=off (no FMA contraction; Terra/LLVM does not
// contract either).  pow(x,2.0) is exact x*x in glibc.

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/mpas_b200.h"

namespace {

struct FieldInfo { const char* name; int entity; int slots; };
const FieldInfo kFields[] = {
#define MPASB200_FIELD(name, entity, slots) {#name, MPASB200_##entity, slots},
#define MPASB200_VFIELD(name) {#name, MPASB200_VERTICAL, 1},
#include "../include/mpas_b200_fields.def"
#undef MPASB200_FIELD
#undef MPASB200_VFIELD
};

struct F3 {            // scalar 3-D field accessor: f(x, k), k in [-1, L]
  double* p; int LS;
  inline double& operator()(int x, int k) const { return p[(size_t)x * LS + (k + 1)]; }
};
struct F3A {           // array-typed 3-D field accessor: f(x, k, slot)
  double* p; int LS; int S;
  inline double& operator()(int x, int k, int s) const { return p[((size_t)x * LS + (k + 1)) * S + s]; }
};

struct Oracle {
  MpasDims d; MpasConfig c;
  int nCells, nEdges, nVertices, L, LS;
  int maxEdges, maxEdges2, vertexDegree, nAdv;
  std::vector<std::vector<double>> store;   // per field id
  // static data, [N+1] rows, resolved ids
  std::vector<int> nEdgesOnCell, edgesOnCell, verticesOnCell, kiteForCell, bdyMaskCell;
  std::vector<double> edgesOnCellSign, edgesOnCell_sign, invAreaCell, latCell, defc_a, defc_b, specZoneMaskCell, lonCell, coeffs_reconstruct;
  std::vector<uint8_t> isShared, inCpr;
  std::vector<uint8_t> cellClass, edgeClass;      // launch classes of MpasMeshPtrs (all 0 when absent)
  int onlyClass[2] = {-1, -1};                   // oracle_set_class: restrict the acoustic step (cells) / divergence damping (edges)
  bool cell_on(int c) const { return onlyClass[0] < 0 || cellClass[c] == onlyClass[0]; }
  bool edge_on(int e) const { return onlyClass[1] < 0 || edgeClass[e] == onlyClass[1]; }
  std::vector<int> cellsOnEdge, verticesOnEdge, nEdgesOnEdge, edgesOnEdge_ECP, edgesOnEdge, nAdvCellsForEdge, advCellsForEdge;
  std::vector<double> weightsOnEdge, dcEdge, dvEdge, invDcEdge, invDvEdge, angleEdge, latEdge, adv_coefs, adv_coefs_3rd,
      meshScalingDel2, meshScalingDel4, specZoneMaskEdge;
  std::vector<int> edgesOnVertex;
  std::vector<double> edgesOnVertexSign, edgesOnVertex_sign, kiteAreasOnVertex, fVertex, invAreaTriangle;
  bool mesh_ok = false;
  int threads = 1;

  int count(int entity) const { return entity == MPASB200_CELL ? nCells : entity == MPASB200_EDGE ? nEdges : entity == MPASB200_VERTEX ? nVertices : 1; }
  F3 f(int id) { return F3{store[id].data(), LS}; }
  F3A fa(int id) { return F3A{store[id].data(), LS, kFields[id].slots}; }
  double* vf(int id) { return store[id].data() + 1; }   // vertical field, index k in [-1, L]
};

#define CF(name) F3 name = o->f(MPASB200_F_##name)
#define VF(name) double* name = o->vf(MPASB200_F_##name)

inline double flux4(double q_im2, double q_im1, double q_i, double q_ip1, double ua) {        // dynamics_tasks.rg:781-783
  return ua * (7. * (q_i + q_im1) - (q_ip1 + q_im2)) / 12.0;
}
inline double flux3(double q_im2, double q_im1, double q_i, double q_ip1, double ua, double coef3) {   // :786-789
  return flux4(q_im2, q_im1, q_i, q_ip1, ua) + coef3 * fabs(ua) * ((q_ip1 - q_im2) - 3. * (q_i - q_im1)) / 12.0;
}

template <class T> void fill_rows(std::vector<T>& dst, const T* src, int n, int w) {
  dst.assign((size_t)(n + 1) * w, T(0));
  if (src) std::memcpy(dst.data(), src, sizeof(T) * (size_t)n * w);
}
void fill_ids(std::vector<int>& dst, const int32_t* src, int n, int w, int target_n, int policy) {
  dst.assign((size_t)(n + 1) * w, target_n);
  if (!src) { std::fill(dst.begin(), dst.end(), policy == MPASB200_INDEX_LITERAL ? 0 : target_n); for (int j = 0; j < w; ++j) dst[(size_t)n * w + j] = target_n; return; }
  for (size_t i = 0; i < (size_t)n * w; ++i) {
    long id = src[i];
    long idx = (policy == MPASB200_INDEX_LITERAL) ? id : id - 1;
    if (idx < 0 || idx > target_n) idx = target_n;
    dst[i] = (int)idx;
  }
  // the pad entity's own connectivity points at pads
}

}  // namespace

#define OMP_FOR _Pragma("omp parallel for schedule(static) num_threads(o->threads)")

extern "C" {

typedef struct Oracle oracle_t;

int oracle_create(const MpasDims* dims, const MpasConfig* cfg, oracle_t** out) {
  if (!dims || !cfg || !out) return MPASB200_EINVAL;
  Oracle* o = new Oracle();
  o->d = *dims; o->c = *cfg;
  o->nCells = dims->nCells; o->nEdges = dims->nEdges; o->nVertices = dims->nVertices; o->L = dims->nVertLevels;
  o->LS = o->L + 2;
  o->maxEdges = dims->maxEdges; o->maxEdges2 = dims->maxEdges2; o->vertexDegree = dims->vertexDegree; o->nAdv = dims->nAdvCells;
  o->store.resize(MPASB200_F_COUNT);
  for (int id = 0; id < MPASB200_F_COUNT; ++id) {
    size_t n = (size_t)(o->count(kFields[id].entity) + 1) * o->LS * kFields[id].slots;
    if (kFields[id].entity == MPASB200_VERTICAL) n = o->LS;
    o->store[id].assign(n, 0.0);
  }
  *out = o;
  return 0;
}
int oracle_destroy(oracle_t* o) { delete o; return 0; }
// the counterpart of mpasb200_set_range(class_range(cls)): restrict the acoustic step to the cells, the divergence damping
// (and the CORRECTED acoustic edge update) to the edges of one launch class; cls < 0 = everything
int oracle_set_class(oracle_t* o, int entity, int cls) { if (entity < 0 || entity > 1) return -1; o->onlyClass[entity] = cls; return 0; }
int oracle_set_threads(oracle_t* o, int n) { o->threads = n < 1 ? 1 : n; return 0; }
int oracle_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int oracle_upload_mesh(oracle_t* o, const MpasMeshPtrs* m) {
  if (!o || !m) return MPASB200_EINVAL;
  const int nC = o->nCells, nE = o->nEdges, nV = o->nVertices, pol = o->c.index_policy;
  const int ME = o->maxEdges, ME2 = o->maxEdges2, VD = o->vertexDegree, NA = o->nAdv;
  fill_rows(o->nEdgesOnCell, m->nEdgesOnCell, nC, 1);
  fill_ids(o->edgesOnCell, m->edgesOnCell, nC, ME, nE, pol);
  fill_ids(o->verticesOnCell, m->verticesOnCell, nC, ME, nV, pol);
  fill_rows(o->kiteForCell, m->kiteForCell, nC, ME);
  fill_rows(o->edgesOnCellSign, m->edgesOnCellSign, nC, ME);
  fill_rows(o->edgesOnCell_sign, m->edgesOnCell_sign, nC, ME);
  fill_rows(o->invAreaCell, m->invAreaCell, nC, 1);
  fill_rows(o->latCell, m->latCell, nC, 1);
  fill_rows(o->defc_a, m->defc_a, nC, ME);
  fill_rows(o->defc_b, m->defc_b, nC, ME);
  fill_rows(o->bdyMaskCell, m->bdyMaskCell, nC, 1);
  fill_rows(o->specZoneMaskCell, m->specZoneMaskCell, nC, 1);
  fill_rows(o->isShared, m->isShared, nC, 1);
  fill_rows(o->cellClass, m->cellClass, nC, 1);
  fill_rows(o->edgeClass, m->edgeClass, nE, 1);
  o->inCpr.assign(nC + 1, 1); o->inCpr[nC] = 0;
  if (m->inCpr) std::memcpy(o->inCpr.data(), m->inCpr, nC);
  fill_ids(o->cellsOnEdge, m->cellsOnEdge, nE, 2, nC, pol);
  fill_ids(o->verticesOnEdge, m->verticesOnEdge, nE, 2, nV, pol);
  fill_rows(o->nEdgesOnEdge, m->nEdgesOnEdge, nE, 1);
  fill_ids(o->edgesOnEdge_ECP, m->edgesOnEdge_ECP, nE, ME2, nE, pol);
  fill_ids(o->edgesOnEdge, m->edgesOnEdge, nE, ME2, nE, pol);
  fill_rows(o->weightsOnEdge, m->weightsOnEdge, nE, ME2);
  fill_rows(o->dcEdge, m->dcEdge, nE, 1);
  fill_rows(o->dvEdge, m->dvEdge, nE, 1);
  fill_rows(o->invDcEdge, m->invDcEdge, nE, 1);
  fill_rows(o->invDvEdge, m->invDvEdge, nE, 1);
  fill_rows(o->angleEdge, m->angleEdge, nE, 1);
  fill_rows(o->latEdge, m->latEdge, nE, 1);
  fill_rows(o->nAdvCellsForEdge, m->nAdvCellsForEdge, nE, 1);
  fill_ids(o->advCellsForEdge, m->advCellsForEdge, nE, NA, nC, pol);
  fill_rows(o->adv_coefs, m->adv_coefs, nE, NA);
  fill_rows(o->adv_coefs_3rd, m->adv_coefs_3rd, nE, NA);
  fill_rows(o->meshScalingDel2, m->meshScalingDel2, nE, 1);
  fill_rows(o->meshScalingDel4, m->meshScalingDel4, nE, 1);
  fill_rows(o->specZoneMaskEdge, m->specZoneMaskEdge, nE, 1);
  fill_ids(o->edgesOnVertex, m->edgesOnVertex, nV, VD, nE, pol);
  fill_rows(o->edgesOnVertexSign, m->edgesOnVertexSign, nV, VD);
  fill_rows(o->edgesOnVertex_sign, m->edgesOnVertex_sign, nV, VD);
  fill_rows(o->kiteAreasOnVertex, m->kiteAreasOnVertex, nV, VD);
  fill_rows(o->fVertex, m->fVertex, nV, 1);
  fill_rows(o->invAreaTriangle, m->invAreaTriangle, nV, 1);
  fill_rows(o->lonCell, m->lonCell, nC, 1);
  fill_rows(o->coeffs_reconstruct, m->coeffs_reconstruct, nC, ME * 3);
  o->mesh_ok = true;
  return 0;
}

// host array layout on this boundary: [n][L+1][slots] contiguous doubles (levels 0..L)
int oracle_upload_field(oracle_t* o, int id, const double* src) {
  if (!o || id < 0 || id >= MPASB200_F_COUNT || !src) return MPASB200_EINVAL;
  const int L1 = o->L + 1, S = kFields[id].slots;
  if (kFields[id].entity == MPASB200_VERTICAL) { std::memcpy(o->store[id].data() + 1, src, sizeof(double) * L1); return 0; }
  const int n = o->count(kFields[id].entity);
  double* dst = o->store[id].data();
  for (int x = 0; x < n; ++x)
    std::memcpy(dst + ((size_t)x * o->LS + 1) * S, src + (size_t)x * L1 * S, sizeof(double) * L1 * S);
  return 0;
}
int oracle_download_field(oracle_t* o, int id, double* dst) {
  if (!o || id < 0 || id >= MPASB200_F_COUNT || !dst) return MPASB200_EINVAL;
  const int L1 = o->L + 1, S = kFields[id].slots;
  if (kFields[id].entity == MPASB200_VERTICAL) { std::memcpy(dst, o->store[id].data() + 1, sizeof(double) * L1); return 0; }
  const int n = o->count(kFields[id].entity);
  const double* src = o->store[id].data();
  for (int x = 0; x < n; ++x)
    std::memcpy(dst + (size_t)x * L1 * S, src + ((size_t)x * o->LS + 1) * S, sizeof(double) * L1 * S);
  return 0;
}
// value of the pad entity (index N) -- only atm_recover_large_step_variables ever writes it
int oracle_download_pad(oracle_t* o, int id, double* dst) {
  const int L1 = o->L + 1, S = kFields[id].slots;
  const int n = o->count(kFields[id].entity);
  std::memcpy(dst, o->store[id].data() + ((size_t)n * o->LS + 1) * S, sizeof(double) * L1 * S);
  return 0;
}

// ------------------------------------------------------------------------------------------
// atm_rk_integration_setup -- dynamics_tasks.rg:747-778
int oracle_rk_integration_setup(oracle_t* o) {
  const int L = o->L;
  CF(ru); CF(ru_save); CF(u); CF(u_2);
  CF(rw); CF(rw_save); CF(rtheta_p); CF(rtheta_p_save); CF(rho_p); CF(rho_p_save);
  CF(w); CF(w_2); CF(theta_m); CF(theta_m_2); CF(rho_zz); CF(rho_zz_2); CF(rho_zz_old_split);
  OMP_FOR
  for (int e = 0; e < o->nEdges; ++e) for (int k = 0; k < L; ++k) { ru_save(e, k) = ru(e, k); u_2(e, k) = u(e, k); }
  OMP_FOR
  for (int c = 0; c < o->nCells; ++c) for (int k = 0; k < L; ++k) {
    rw_save(c, k) = rw(c, k); rtheta_p_save(c, k) = rtheta_p(c, k); rho_p_save(c, k) = rho_p(c, k);
    w_2(c, k) = w(c, k); theta_m_2(c, k) = theta_m(c, k); rho_zz_2(c, k) = rho_zz(c, k); rho_zz_old_split(c, k) = rho_zz(c, k);
  }
  if (o->c.config_scalar_advection) {      // MPAS: scalars_2 = scalars_1 (the reference's setup has no scalar line: it never advects)
    F3A scalars = o->fa(MPASB200_F_scalars), scalars_old = o->fa(MPASB200_F_scalars_old);
    OMP_FOR
    for (int c = 0; c < o->nCells; ++c) for (int k = 0; k < L; ++k) for (int s = 0; s < scalars.S; ++s) scalars_old(c, k, s) = scalars(c, k, s);
  }
  return 0;
}

// atm_advance_scalars -- ABSENT from the reference (rk_timestep.rg:465,469,485 are "SKIPPING" comments; the storage is
// cell_fs.scalars, data_structures.rg:36).  PARITY UNPINNED.  This restates atm_advance_scalars_work of MPAS-Atmosphere v7.0
// (src/core_atmosphere/dynamics/mpas_atm_time_integration.F; upstream, not under /root/reference), loop by loop, non-monotonic
// branch, scalar_tend_save = 0 (no physics), advance_density = false, in the reference's 0-based level convention:
//   edges : horiz_flux_arr(s,k,e) = sum_j (adv_coefs(j,e) + sign(1, uhAvg(k,e)) * adv_coefs_3rd(j,e)) * scalar_new(s,k,advCellsForEdge(j,e))
//   cells : tend(s,k) = -sum_i edgesOnCell_sign(i,c) * uhAvg(k,e_i) * horiz_flux_arr(s,k,e_i);  tend = tend * invAreaCell
//           wdtn(s,0) = wdtn(s,L) = 0; wdtn(s,1) and wdtn(s,L-1) second order with fnm/fnp; flux3 with coef_3rd_order between
//           scalar_new(s,k,c) = (scalar_old(s,k,c) * rho_zz_old(k,c) + dt * (tend(s,k) - rdnw(k) * (wdtn(s,k+1) - wdtn(s,k)))) / rho_zz_new(k,c)
// Bindings: uhAvg = er.ruAvg, wwAvg = cr.wwAvg (data_structures.rg:184,514: "used in scalar transport"), fnm/fnp/rdnw =
// vert_r.fzm/fzp/rdzw, edgesOnCell_sign = the field atm_compute_signs fills (cr.edgesOnCellSign, dynamics_tasks.rg:74-86),
// scalar_new = cr.scalars, scalar_old = scalars_old (saved by the setup task), rho_zz_new = cr.rho_zz, rho_zz_old =
// cr.rho_zz_old_split (the setup task's copy of rho_zz at the start of the step).
int oracle_advance_scalars(oracle_t* o, double dt, int rk_step) {
  (void)rk_step;
  const int L = o->L, nC = o->nCells, nE = o->nEdges, ME = o->maxEdges, NA = o->nAdv;
  const double coef_3rd_order = o->c.config_coef_3rd_order;
  F3A scalars = o->fa(MPASB200_F_scalars), scalars_old = o->fa(MPASB200_F_scalars_old);
  const int NS = scalars.S;
  CF(ruAvg); CF(wwAvg); CF(rho_zz); CF(rho_zz_old_split);
  VF(fzm); VF(fzp); VF(rdzw);
  std::vector<double> flux((size_t)(nE + 1) * L * NS, 0.0);               // horiz_flux_arr; the pad edge carries no flux
  OMP_FOR
  for (int e = 0; e < nE; ++e)
    for (int j = 0; j < o->nAdvCellsForEdge[e]; ++j) {
      const int iAdvCell = o->advCellsForEdge[e * NA + j];
      for (int k = 0; k < L; ++k) {
        const double scalar_weight = o->adv_coefs[e * NA + j] + copysign(1.0, ruAvg(e, k)) * o->adv_coefs_3rd[e * NA + j];
        for (int s = 0; s < NS; ++s) flux[((size_t)e * L + k) * NS + s] += scalar_weight * scalars(iAdvCell, k, s);
      }
    }
  OMP_FOR
  for (int c = 0; c < nC; ++c) {
    std::vector<double> tend((size_t)L * NS, 0.0), wdtn((size_t)(L + 1) * NS, 0.0);
    for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
      const int e = o->edgesOnCell[c * ME + i];
      for (int k = 0; k < L; ++k)
        for (int s = 0; s < NS; ++s)
          tend[(size_t)k * NS + s] -= o->edgesOnCellSign[c * ME + i] * ruAvg(e, k) * flux[((size_t)e * L + k) * NS + s];
    }
    for (int k = 0; k < L; ++k) for (int s = 0; s < NS; ++s) tend[(size_t)k * NS + s] = tend[(size_t)k * NS + s] * o->invAreaCell[c];
    for (int s = 0; s < NS; ++s) {
      if (L > 1) wdtn[(size_t)1 * NS + s] = wwAvg(c, 1) * (fzm[1] * scalars(c, 1, s) + fzp[1] * scalars(c, 0, s));
      for (int k = 2; k < L - 1; ++k)
        wdtn[(size_t)k * NS + s] = flux3(scalars(c, k - 2, s), scalars(c, k - 1, s), scalars(c, k, s), scalars(c, k + 1, s), wwAvg(c, k), coef_3rd_order);
      if (L > 2) wdtn[(size_t)(L - 1) * NS + s] = wwAvg(c, L - 1) * (fzm[L - 1] * scalars(c, L - 1, s) + fzp[L - 1] * scalars(c, L - 2, s));
    }
    for (int k = 0; k < L; ++k)
      for (int s = 0; s < NS; ++s)
        scalars(c, k, s) = (scalars_old(c, k, s) * rho_zz_old_split(c, k)
                            + dt * (tend[(size_t)k * NS + s] - rdzw[k] * (wdtn[(size_t)(k + 1) * NS + s] - wdtn[(size_t)k * NS + s]))) / rho_zz(c, k);
  }
  return 0;
}

// atm_init_coupled_diagnostics -- dynamics_tasks.rg:651-725 (one-time task of atm_core_init; levels 0..nVertLevels-1)
int oracle_init_coupled_diagnostics(oracle_t* o) {
  const int L = o->L, nC = o->nCells, nE = o->nEdges, ME = o->maxEdges;
  const double rgas = o->c.rgas, rcv = rgas / (o->c.cp - rgas);
  const int p0 = 100000;                                                        // an integer in the reference (:668)
  CF(rho_zz); CF(zz); CF(ru); CF(u); CF(rw); CF(w); CF(rho_p); CF(rho_base); CF(rtheta_base); CF(theta_base); CF(rtheta_p); CF(theta_m);
  CF(exner); CF(exner_base); CF(pressure_p); CF(pressure_base);
  F3A zb_cell = o->fa(MPASB200_F_zb_cell), zb3_cell = o->fa(MPASB200_F_zb3_cell);
  VF(fzm); VF(fzp);
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) rho_zz(c, k) /= zz(c, k);                                    // :676
  OMP_FOR
  for (int e = 0; e < nE; ++e) {
    const int cell1 = o->cellsOnEdge[e * 2], cell2 = o->cellsOnEdge[e * 2 + 1];
    for (int k = 0; k < L; ++k) ru(e, k) = 0.5 * u(e, k) * (rho_zz(cell1, k) + rho_zz(cell2, k));                        // :679-683
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) {
    for (int k = 0; k < L; ++k) {
      rw(c, k) = 0;
      if (k > 0 && k < L)                                                                                               // :691-694
        rw(c, k) = w(c, k) * (fzp[k] * rho_zz(c, k - 1) + fzm[k] * rho_zz(c, k)) * (fzp[k] * zz(c, k - 1) + fzm[k] * zz(c, k));
    }
    for (int k = 0; k < L; ++k)                                                                                          // :698-709
      for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
        const int e = o->edgesOnCell[c * ME + i];
        if (k > 0) {
          const double flux = fzm[k] * ru(e, k) + fzp[k] * ru(e, k - 1);
          rw(c, k) -= o->edgesOnCellSign[c * ME + i] * (zb_cell(c, k, i) + copysign(1.0, flux) * zb3_cell(c, k, i)) * flux
                      * (fzp[k] * zz(c, k - 1) + fzm[k] * zz(c, k));
        }
      }
    for (int k = 0; k < L; ++k) {                                                                                        // :711-718
      rho_p(c, k) = rho_zz(c, k) - rho_base(c, k);
      rtheta_base(c, k) = theta_base(c, k) * rho_base(c, k);
      rtheta_p(c, k) = theta_m(c, k) * rho_p(c, k) + rho_base(c, k) * (theta_m(c, k) - theta_base(c, k));
      exner(c, k) = pow(zz(c, k) * (rgas / p0) * (rtheta_p(c, k) + rtheta_base(c, k)), rcv);
      exner_base(c, k) = pow(zz(c, k) * (rgas / p0) * (rtheta_base(c, k)), rcv);
      pressure_p(c, k) = zz(c, k) * rgas * (exner(c, k) * rtheta_p(c, k) + rtheta_base(c, k) * (exner(c, k) - exner_base(c, k)));
      pressure_base(c, k) = zz(c, k) * rgas * exner_base(c, k) * rtheta_base(c, k);
    }
  }
  return 0;
}

// mpas_reconstruct_2d -- dynamics_tasks.rg:1894-1948
int oracle_reconstruct_2d(oracle_t* o, int includeHalos, int on_a_sphere) {
  (void)includeHalos;                                                           // nCellsReconstruct = nCells either way (:1909-1912)
  const int L = o->L, nC = o->nCells, ME = o->maxEdges;
  CF(u); CF(uReconstructX); CF(uReconstructY); CF(uReconstructZ); CF(uReconstructZonal); CF(uReconstructMeridional);
  OMP_FOR
  for (int c = 0; c < nC; ++c) {
    for (int k = 0; k < L; ++k) {
      uReconstructX(c, k) = 0.0; uReconstructY(c, k) = 0.0; uReconstructZ(c, k) = 0.0;
      for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
        const int e = o->edgesOnCell[c * ME + i];
        const double* cf = &o->coeffs_reconstruct[((size_t)c * ME + i) * 3];
        uReconstructX(c, k) += cf[0] * u(e, k);
        uReconstructY(c, k) += cf[1] * u(e, k);
        uReconstructZ(c, k) += cf[2] * u(e, k);
      }
    }
    if (on_a_sphere) {
      const double clat = cos(o->latCell[c]), slat = sin(o->latCell[c]), clon = cos(o->lonCell[c]), slon = sin(o->lonCell[c]);
      for (int k = 0; k < L; ++k) {
        uReconstructZonal(c, k) = -uReconstructX(c, k) * slon + uReconstructY(c, k) * clon;
        uReconstructMeridional(c, k) = -(uReconstructX(c, k) * clon + uReconstructY(c, k) * slon) * slat + uReconstructZ(c, k) * clat;
      }
    } else {
      for (int k = 0; k < L; ++k) { uReconstructZonal(c, k) = uReconstructX(c, k); uReconstructMeridional(c, k) = uReconstructY(c, k); }
    }
  }
  return 0;
}

// atm_compute_moist_coefficients -- dynamics_tasks.rg:460-502  (edge loop is empty: cqu never written)
// ---- the mesh-only producers of atm_core_init --------------------------------------------------------------
// Literal loops over the RAW stored ids in the caller's numbering.  `R(id, n)` is rule M2 (the pad entity n reads as zero).
namespace {
struct RawMesh {
  const MpasInitMesh* m; int nC, nE, nV, ME, VD, pol;
  long R(long id, int n) const { long idx = (pol == MPASB200_INDEX_LITERAL) ? id : id - 1; if (idx < 0 || idx > n) idx = n; return idx; }
  int off() const { return pol == MPASB200_INDEX_LITERAL ? 0 : 1; }       // what a 0-based loop index is offset by to equal a stored id
  int nEdgesOnCell(long c) const { return c < nC ? m->nEdgesOnCell[c] : 0; }
  int cellsOnCell(long c, int i) const { return c < nC ? m->cellsOnCell[c * ME + i] : 0; }
  int cellsOnEdge(long e, int i) const { return e < nE ? m->cellsOnEdge[e * 2 + i] : 0; }
  int verticesOnEdge(long e, int i) const { return e < nE ? m->verticesOnEdge[e * 2 + i] : 0; }
  int cellsOnVertex(long v, int i) const { return v < nV ? m->cellsOnVertex[v * VD + i] : 0; }
};
RawMesh raw_of(const Oracle* o, const MpasInitMesh* m) { return RawMesh{m, o->nCells, o->nEdges, o->nVertices, o->maxEdges, o->vertexDegree, o->c.index_policy}; }
}  // namespace

// atm_compute_signs, level-0 part   dynamics_tasks.rg:60-86, 113-129
int oracle_compute_signs(oracle_t* o, const MpasInitMesh* m, double* edgesOnVertexSign, double* edgesOnCellSign, int32_t* kiteForCell) {
  if (!o || !m || !m->edgesOnVertex || !m->verticesOnEdge || !m->nEdgesOnCell || !m->edgesOnCell || !m->cellsOnEdge || !m->verticesOnCell || !m->cellsOnVertex)
    return MPASB200_EINVAL;
  const RawMesh r = raw_of(o, m);
  const int nC = r.nC, nE = r.nE, nV = r.nV, ME = r.ME, VD = r.VD;
  if (edgesOnVertexSign)
    for (int iVtx = 0; iVtx < nV; ++iVtx)                                                          // :60-72
      for (int i = 0; i < VD; ++i) {
        const int e = m->edgesOnVertex[(size_t)iVtx * VD + i];
        if (e <= nE) edgesOnVertexSign[(size_t)iVtx * VD + i] = (iVtx + r.off() == r.verticesOnEdge(r.R(e, nE), 1)) ? 1.0 : -1.0;
        else edgesOnVertexSign[(size_t)iVtx * VD + i] = 0.0;
      }
  if (edgesOnCellSign)
    for (int iCell = 0; iCell < nC; ++iCell) {                                                     // :74-86
      for (int i = 0; i < ME; ++i) edgesOnCellSign[(size_t)iCell * ME + i] = 0.0;                  // rule M1: never written = 0
      for (int i = 0; i < m->nEdgesOnCell[iCell] && i < ME; ++i) {
        const int e = m->edgesOnCell[(size_t)iCell * ME + i];
        if (e <= nE) edgesOnCellSign[(size_t)iCell * ME + i] = (iCell + r.off() == r.cellsOnEdge(r.R(e, nE), 0)) ? 1.0 : -1.0;
        else edgesOnCellSign[(size_t)iCell * ME + i] = 0.0;
      }
    }
  if (kiteForCell)
    for (int iCell = 0; iCell < nC; ++iCell) {                                                     // :113-129
      for (int i = 0; i < ME; ++i) kiteForCell[(size_t)iCell * ME + i] = 0;
      for (int i = 0; i < m->nEdgesOnCell[iCell] && i < ME; ++i) {
        const int iVtx = m->verticesOnCell[(size_t)iCell * ME + i];
        if (iVtx <= nV) {
          for (int j = 1; j < VD; ++j)
            if (iCell + r.off() == r.cellsOnVertex(r.R(iVtx, nV), j)) { kiteForCell[(size_t)iCell * ME + i] = j; break; }
        } else kiteForCell[(size_t)iCell * ME + i] = 1;
      }
    }
  return 0;
}

// atm_compute_signs, 3-D part   dynamics_tasks.rg:88-110 (resolved numbering: `iCell.x == cellsOnEdge[0]` is `iCell == resolved cell1`
// under either index policy; an edge id that resolves to the pad copies the pad's zeros)
int oracle_compute_zb_cell(oracle_t* o) {
  if (!o->mesh_ok) return MPASB200_ESTATE;
  const int nC = o->nCells, L = o->L, ME = o->maxEdges;
  F3A zb = o->fa(MPASB200_F_zb), zb3 = o->fa(MPASB200_F_zb3), zb_cell = o->fa(MPASB200_F_zb_cell), zb3_cell = o->fa(MPASB200_F_zb3_cell);
  OMP_FOR
  for (int iCell = 0; iCell < nC; ++iCell)
    for (int k = 0; k <= L; ++k)
      for (int i = 0; i < o->nEdgesOnCell[iCell]; ++i) {
        const int e = o->edgesOnCell[(size_t)iCell * ME + i];
        const int side = (iCell == o->cellsOnEdge[(size_t)e * 2]) ? 0 : 1;
        zb_cell(iCell, k, i) = zb(e, k, side);
        zb3_cell(iCell, k, i) = zb3(e, k, side);
      }
  return 0;
}

// atm_adv_coef_compression   dynamics_tasks.rg:133-269
int oracle_adv_coef_compression(oracle_t* o, const MpasInitMesh* m, int32_t* nAdvCellsForEdge, int32_t* advCellsForEdge,
                                double* adv_coefs, double* adv_coefs_3rd) {
  if (!o || !m || !m->cellsOnEdge || !m->cellsOnCell || !m->nEdgesOnCell || !m->dcEdge || !m->dvEdge || !nAdvCellsForEdge || !advCellsForEdge || !adv_coefs || !adv_coefs_3rd)
    return MPASB200_EINVAL;
  const RawMesh r = raw_of(o, m);
  const int nC = r.nC, nE = r.nE, ME = r.ME, NA = o->nAdv;
  const int W = 2 + 2 * ME;
  std::vector<int> cell_list(W);
  std::vector<double> a(W), a3(W);
  for (int iEdge = 0; iEdge < nE; ++iEdge) {
    nAdvCellsForEdge[iEdge] = 0;
    for (int j = 0; j < NA; ++j) { advCellsForEdge[(size_t)iEdge * NA + j] = 0; adv_coefs[(size_t)iEdge * NA + j] = 0.0; adv_coefs_3rd[(size_t)iEdge * NA + j] = 0.0; }
    const int cell1 = m->cellsOnEdge[(size_t)iEdge * 2], cell2 = m->cellsOnEdge[(size_t)iEdge * 2 + 1];
    if (!(cell1 <= nC || cell2 <= nC)) continue;
    const long i1 = r.R(cell1, nC), i2 = r.R(cell2, nC);
    auto d2 = [&](int idx) { return (m->deriv_two && idx < 2 * NA) ? m->deriv_two[(size_t)iEdge * 2 * NA + idx] : 0.0; };   // past the array: 0
    cell_list[0] = cell1; cell_list[1] = cell2;
    int n = 1;
    for (int i = 0; i < r.nEdgesOnCell(i1); ++i)
      if (r.cellsOnCell(i1, i) != cell2) { n += 1; cell_list[n] = r.cellsOnCell(i1, i); }
    for (int iCell = 0; iCell < r.nEdgesOnCell(i2); ++iCell) {
      bool addcell = true;
      for (int i = 0; i < n; ++i) if (cell_list[i] == r.cellsOnCell(i2, iCell)) addcell = false;
      if (addcell && n < ME - 1) { n += 1; cell_list[n] = r.cellsOnCell(i2, iCell); }
    }
    nAdvCellsForEdge[iEdge] = n;
    for (int iCell = 0; iCell < n && iCell < NA; ++iCell) advCellsForEdge[(size_t)iEdge * NA + iCell] = cell_list[iCell];
    std::fill(a.begin(), a.end(), 0.0); std::fill(a3.begin(), a3.end(), 0.0);
    auto j_in_of = [&](int target) { int j_in = 0; for (int j = 0; j < n; ++j) if (cell_list[j] == target) j_in = j; return j_in; };
    int j_in = j_in_of(cell1);
    a[j_in] += d2(0); a3[j_in] += d2(0);
    for (int iCell = 0; iCell < r.nEdgesOnCell(i1); ++iCell) {
      j_in = j_in_of(r.cellsOnCell(i1, iCell));
      a[j_in] += d2(iCell * NA + 0); a3[j_in] += d2(iCell * NA + 0);
    }
    j_in = j_in_of(cell2);
    a[j_in] += d2(1); a3[j_in] += d2(1);
    for (int iCell = 0; iCell < r.nEdgesOnCell(i2); ++iCell) {
      j_in = j_in_of(r.cellsOnCell(i2, iCell));
      a[j_in] += d2(iCell * NA + 1); a3[j_in] += d2(iCell * NA + 1);
    }
    const double dc = m->dcEdge[iEdge], dv = m->dvEdge[iEdge];
    for (int j = 0; j < n; ++j) {
      a[j] = -1.0 * pow(dc, 2) * a[j] / 12;
      a3[j] = -1.0 * pow(dc, 2) * a3[j] / 12;
    }
    a[j_in_of(cell1)] += 0.5;
    a[j_in_of(cell2)] += 0.5;
    for (int j = 0; j < n; ++j) { a[j] *= dv; a3[j] *= dv; }
    for (int j = 0; j < NA && j < W; ++j) { adv_coefs[(size_t)iEdge * NA + j] = a[j]; adv_coefs_3rd[(size_t)iEdge * NA + j] = a3[j]; }
  }
  return 0;
}

// atm_compute_mesh_scaling   dynamics_tasks.rg:595-646 (del2 / del4 factors)
int oracle_compute_mesh_scaling(oracle_t* o, const MpasInitMesh* m, const double* meshDensity, int scale_with_mesh, double* del2, double* del4) {
  if (!o || !m || !m->cellsOnEdge || !del2 || !del4) return MPASB200_EINVAL;
  const RawMesh r = raw_of(o, m);
  for (int iEdge = 0; iEdge < r.nE; ++iEdge) { del2[iEdge] = 1.0; del4[iEdge] = 1.0; }
  if (scale_with_mesh)
    for (int iEdge = 0; iEdge < r.nE; ++iEdge) {
      const long c1 = r.R(m->cellsOnEdge[(size_t)iEdge * 2], r.nC), c2 = r.R(m->cellsOnEdge[(size_t)iEdge * 2 + 1], r.nC);
      const double md = ((c1 < r.nC ? meshDensity[c1] : 0.0) + (c2 < r.nC ? meshDensity[c2] : 0.0)) / 2.0;
      del2[iEdge] = 1.0 / pow(md, 0.25);
      del4[iEdge] = 1.0 / pow(md, 0.75);
    }
  return 0;
}

// atm_compute_damping_coefs   dynamics_tasks.rg:274-300
int oracle_compute_damping_coefs(oracle_t* o, const double* meshDensity, double config_zd, double config_xnutr) {
  if (!o || !meshDensity) return MPASB200_EINVAL;
  F3 zgrid = o->f(MPASB200_F_zgrid), dss = o->f(MPASB200_F_dss);
  const double m1 = -1.0, pii = acos(m1), dx_scale_power = 1.0;
  for (int iCell = 0; iCell < o->nCells; ++iCell)
    for (int k = 0; k < o->L; ++k) {
      dss(iCell, k) = 0.0;
      const double zt = zgrid(iCell, o->L);
      const double z = 0.5 * (zgrid(iCell, k) + zgrid(iCell, k + 1));
      if (z > config_zd) {
        dss(iCell, k) = config_xnutr * pow(sin(0.5 * pii * (z - config_zd) / (zt - config_zd)), 2.0);
        dss(iCell, k) /= pow(meshDensity[iCell], (0.25 * dx_scale_power));
      }
    }
  return 0;
}

// atm_couple_coef_3rd_order   dynamics_tasks.rg:303-325 (zb3_cell: `cr[{iCell, 0}]` is LEVEL 0 only)
int oracle_couple_coef_3rd_order(oracle_t* o, double coef, double* adv_coefs_3rd) {
  if (adv_coefs_3rd) for (size_t i = 0; i < (size_t)o->nEdges * o->nAdv; ++i) adv_coefs_3rd[i] *= coef;
  F3A zb3_cell = o->fa(MPASB200_F_zb3_cell);
  for (int iCell = 0; iCell < o->nCells; ++iCell) for (int j = 0; j < o->maxEdges; ++j) zb3_cell(iCell, 0, j) *= coef;
  return 0;
}

int oracle_compute_moist_coefficients(oracle_t* o) {
  const int L = o->L;
  CF(qtot); CF(cqw);
  OMP_FOR
  for (int c = 0; c < o->nCells; ++c) for (int k = 0; k < L; ++k) qtot(c, k) = 0.0;            // :473-482
  OMP_FOR
  for (int c = 0; c < o->nCells; ++c) for (int k = 0; k < L; ++k)
    if (k > 0) { double qtotal = 0.5 * (qtot(c, k) + qtot(c, k - 1)); cqw(c, k) = 1.0 / (1.0 + qtotal); }   // :484-489
  return 0;
}

// atm_compute_vert_imp_coefs -- dynamics_tasks.rg:513-592
int oracle_compute_vert_imp_coefs(oracle_t* o, double dts) {
  const int L = o->L, nC = o->nCells;
  const double dtseps = .5 * dts * (1.0 + o->c.config_epssm);        // :529
  const double rgas = o->c.rgas, rcv = rgas / (o->c.cp - rgas), c2 = o->c.cp * rcv, gravity = o->c.gravity;
  VF(cofrz); VF(rdzw); VF(rdzu); VF(fzm); VF(fzp);
  CF(zz); CF(cqw); CF(exner); CF(theta_m); CF(qtot); CF(rho_base); CF(rtheta_base); CF(rtheta_p); CF(exner_base);
  CF(cofwr); CF(cofwz); CF(coftz); CF(cofwt); CF(a_tri); CF(b_tri); CF(c_tri); CF(alpha_tri); CF(gamma_tri);
  for (int k = 0; k < L; ++k) cofrz[k] = dtseps * rdzw[k];           // :537-539
  for (int c = 0; c < nC; ++c) gamma_tri(c, 0) = 0.0;                // :541-547
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {         // :550-564
    if (k > 0) cofwr(c, k) = .5 * dtseps * gravity * (fzm[k] * zz(c, k) + fzp[k] * zz(c, k - 1));
    coftz(c, k) = 0.0;
    if (k > 0) {
      cofwz(c, k) = dtseps * c2 * (fzm[k] * zz(c, k) + fzp[k] * zz(c, k - 1)) * rdzu[k] * cqw(c, k) * (fzm[k] * exner(c, k) + fzp[k] * exner(c, k - 1));
      coftz(c, k) = dtseps * (fzm[k] * theta_m(c, k) + fzp[k] * theta_m(c, k - 1));
    }
    double qtotal = qtot(c, k);
    cofwt(c, k) = .5 * dtseps * rcv * zz(c, k) * gravity * rho_base(c, k) / (1.0 + qtotal) * exner(c, k) / ((rtheta_base(c, k) + rtheta_p(c, k)) * exner_base(c, k));
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) if (k > 0) {   // :566-578
    a_tri(c, k) = -1.0 * cofwz(c, k) * coftz(c, k - 1) * rdzw[k - 1] * zz(c, k - 1)
                  + cofwr(c, k) * cofrz[k - 1] - cofwt(c, k - 1) * coftz(c, k - 1) * rdzw[k - 1];
    b_tri(c, k) = 1.0 + cofwz(c, k) * (coftz(c, k) * rdzw[k] * zz(c, k) + coftz(c, k) * rdzw[k - 1] * zz(c, k - 1))
                  - coftz(c, k) * (cofwt(c, k) * rdzw[k] - cofwt(c, k) * rdzw[k - 1]) + cofwr(c, k) * ((cofrz[k] - cofrz[k - 1]));
    c_tri(c, k) = -1.0 * cofwz(c, k) * coftz(c, k + 1) * rdzw[k] * zz(c, k)
                  - cofwr(c, k) * cofrz[k] + cofwt(c, k) * coftz(c, k + 1) * rdzw[k];
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) if (k > 0)      // :580-585 (gamma_tri from the PREVIOUS call)
    alpha_tri(c, k) = 1.0 / (b_tri(c, k) - a_tri(c, k) * gamma_tri(c, k - 1));
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) if (k > 0)      // :587-591
    gamma_tri(c, k) = c_tri(c, k) * alpha_tri(c, k);
  return 0;
}

// atm_compute_solve_diagnostics -- dynamics_tasks.rg:328-454
int oracle_compute_solve_diagnostics(oracle_t* o, int hollingsworth, int rk_step) {
  const int L = o->L, nC = o->nCells, nE = o->nEdges, nV = o->nVertices, ME = o->maxEdges, ME2 = o->maxEdges2, VD = o->vertexDegree;
  CF(h); CF(h_edge); CF(ke_edge); CF(u); CF(vorticity); CF(divergence); CF(ke); CF(ke_vertex); CF(v); CF(pv_vertex); CF(pv_edge);
  OMP_FOR
  for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k) {               // :346-353
    int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
    h_edge(e, k) = 0.5 * (h(cell1, k) + h(cell2, k));
    double efac = o->dcEdge[e] * o->dvEdge[e];
    ke_edge(e, k) = efac * pow(u(e, k), 2);
  }
  OMP_FOR
  for (int iv = 0; iv < nV; ++iv) for (int k = 0; k < L; ++k) {            // :356-366
    vorticity(iv, k) = 0.0;
    for (int i = 0; i < VD; ++i) {
      int e = o->edgesOnVertex[iv * VD + i];
      double s = o->edgesOnVertexSign[iv * VD + i] * o->dcEdge[e];
      vorticity(iv, k) += s * u(e, k);
    }
    vorticity(iv, k) *= o->invAreaTriangle[iv];
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {               // :369-379 (s + u, Q6)
    divergence(c, k) = 0.0;
    for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
      int e = o->edgesOnCell[c * ME + i];
      double s = o->edgesOnCellSign[c * ME + i] * o->dvEdge[e];
      divergence(c, k) += s + u(e, k);
    }
    double r = o->invAreaCell[c];
    divergence(c, k) *= r;
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {               // :382-390
    ke(c, k) = 0.0;
    for (int i = 0; i < o->nEdgesOnCell[c]; ++i) { int e = o->edgesOnCell[c * ME + i]; ke(c, k) += 0.25 * ke_edge(e, k); }
    ke(c, k) *= o->invAreaCell[c];
  }
  if (hollingsworth) {                                                      // :392-418
    OMP_FOR
    for (int iv = 0; iv < nV; ++iv) for (int k = 0; k < L; ++k) {
      double r = 0.25 * o->invAreaTriangle[iv];
      ke_vertex(iv, k) = (ke_edge(o->edgesOnVertex[iv * VD + 0], k) + ke_edge(o->edgesOnVertex[iv * VD + 1], k) + ke_edge(o->edgesOnVertex[iv * VD + 2], k)) * r;
    }
    double ke_fact = 1.0 - 0.375;
    OMP_FOR
    for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) ke(c, k) *= ke_fact;
    OMP_FOR
    for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {
      double r = o->invAreaCell[c];
      for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
        int iv = o->verticesOnCell[c * ME + i];
        int j = o->kiteForCell[c * ME + i];
        ke(c, k) += (1.0 - ke_fact) * o->kiteAreasOnVertex[iv * VD + j] * ke_vertex(iv, k) * r;
      }
    }
  }
  bool reconstruct_v = true;                                                // :422-428
  if (rk_step != -1 && rk_step != 2) reconstruct_v = false;
  if (reconstruct_v) {
    OMP_FOR
    for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k) {             // :431-438 (loop from i = 1, Q7)
      v(e, k) = 0;
      for (int i = 1; i < o->nEdgesOnEdge[e]; ++i) {
        int eoe = o->edgesOnEdge_ECP[e * ME2 + i];
        v(e, k) += o->weightsOnEdge[e * ME2 + i] * u(eoe, k);
      }
    }
  }
  OMP_FOR
  for (int iv = 0; iv < nV; ++iv) for (int k = 0; k < L; ++k) pv_vertex(iv, k) = o->fVertex[iv] + vorticity(iv, k);   // :443-445
  OMP_FOR
  for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k)                 // :449-451
    pv_edge(e, k) = 0.5 * (pv_vertex(o->verticesOnEdge[e * 2 + 0], k) + pv_vertex(o->verticesOnEdge[e * 2 + 1], k));
  return 0;
}

// atm_compute_dyn_tend_work -- dynamics_tasks.rg:814-1480
int oracle_compute_dyn_tend(oracle_t* o, int rk_step, double dt, int config_horiz_mixing, double config_mpas_cam_coef,
                            int config_mix_full, int config_rayleigh_damp_u) {
  const int L = o->L, nC = o->nCells, nE = o->nEdges, nV = o->nVertices, ME = o->maxEdges, ME2 = o->maxEdges2, VD = o->vertexDegree, NA = o->nAdv;
  const MpasConfig& C = o->c;
  VF(fzm); VF(fzp); VF(rdzu); VF(rdzw); VF(u_init); VF(v_init);
  CF(cqw); CF(divergence); CF(ke); CF(pressure_p); CF(qtot); CF(rho_base); CF(rho_zz); CF(rho_p_save); CF(rt_diabatic_tend);
  CF(rw); CF(rw_save); CF(t_init); CF(tend_rho_physics); CF(tend_rtheta_physics); CF(theta_m); CF(theta_m_save);
  CF(uReconstructZonal); CF(uReconstructMeridional); CF(w); CF(zgrid); CF(zz);
  CF(cqu); CF(pv_edge); CF(rho_edge); CF(ru); CF(ru_save); CF(tend_ru_physics); CF(u); CF(v); CF(zxu);
  CF(vorticity);
  CF(rthdynten); CF(tend_rho); CF(tend_rtheta_adv);
  CF(delsq_divergence); CF(delsq_theta); CF(delsq_w); CF(dpdz); CF(flux_arr); CF(h_divergence); CF(kdiff); CF(ru_edge_w);
  CF(tend_theta); CF(tend_theta_euler); CF(tend_w_euler); CF(wdtz); CF(wdwz);
  CF(delsq_u); CF(q); CF(tend_u); CF(tend_u_euler); CF(u_mix); CF(wduz);
  CF(delsq_vorticity);

  double prandtl_inv = 1.0 / C.prandtl;                                     // :847
  double invDt = 1.0 / dt;
  double r_earth = C.sphere_radius;
  double inv_r_earth = 1.0 / r_earth;
  double v_mom_eddy_visc2 = C.config_v_mom_eddy_visc2;
  double v_theta_eddy_visc2 = C.config_v_theta_eddy_visc2;
  double h_mom_eddy_visc4 = C.config_h_mom_eddy_visc4;
  double h_theta_eddy_visc4 = C.config_h_theta_eddy_visc4;

  if (rk_step == 0) {                                                       // :858
    if (config_horiz_mixing == MPASB200_MIX_2D_SMAGORINSKY) {              // :861 (Q11: an enum here)
      double c_s = C.config_smagorinsky_coef;
      OMP_FOR
      for (int c = 0; c < nC; ++c) {
        std::vector<double> d_diag(L), d_off_diag(L);
        // the reference recomputes both arrays for every (cell, level) point (Q12, O(L^2));
        // the values do not depend on the point's level, so once per cell is the same numbers.
        for (int k = 0; k < L; ++k) { d_diag[k] = 0.0; d_off_diag[k] = 0.0; }
        for (int iEdge = 0; iEdge < o->nEdgesOnCell[c]; ++iEdge) for (int k = 0; k < L; ++k) {
          int e = o->edgesOnCell[c * ME + iEdge];
          d_diag[k] += o->defc_a[c * ME + iEdge] * u(e, k) - o->defc_b[c * ME + iEdge] * v(e, k);
          d_off_diag[k] += o->defc_b[c * ME + iEdge] * u(e, k) + o->defc_a[c * ME + iEdge] * v(e, k);
        }
        for (int k = 0; k < L; ++k)                                        // :884-886
          kdiff(c, k) = std::min(pow(c_s * C.config_len_disp, 2.0) * sqrt(pow(d_diag[k], 2.0) + pow(d_off_diag[k], 2.0)),
                                 (0.01 * pow(C.config_len_disp, 2.0)) * invDt);
      }
      h_mom_eddy_visc4 = C.config_visc4_2dsmag * pow(C.config_len_disp, 3.0);   // :889
      h_theta_eddy_visc4 = h_mom_eddy_visc4;
    } else if (config_horiz_mixing == MPASB200_MIX_2D_FIXED) {             // :892-896
      OMP_FOR
      for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) kdiff(c, k) = 0.0;
    }
    if (config_mpas_cam_coef > 0.0) {                                       // :898-916 (cell_range stops at L-1)
      OMP_FOR
      for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k)
        if (k >= L - 2 && k <= L) { int p = k - (L - 2); kdiff(c, k) = std::max(kdiff(c, k), pow(2, p) * 2.0833 * C.config_len_disp * config_mpas_cam_coef); }
    }
  }

  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {               // :924-931
    h_divergence(c, k) = 0.0;
    for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
      int e = o->edgesOnCell[c * ME + i];
      double edge_sign = o->edgesOnCell_sign[c * ME + i] * o->dvEdge[e];
      h_divergence(c, k) += edge_sign * ru(e, k);
    }
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) { double r = o->invAreaCell[c]; h_divergence(c, k) *= r; }   // :935-938

  if (rk_step == 0) {                                                       // :942-951
    OMP_FOR
    for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {
      tend_rho(c, k) = -h_divergence(c, k) - rdzw[k] * (rw(c, k + 1) - rw(c, k) + tend_rho_physics(c, k));
      dpdz(c, k) = -C.gravity * (rho_base(c, k) * (qtot(c, k)) + rho_p_save(c, k) * (1.0 + qtot(c, k)));
    }
  }

  // -------- U section
  OMP_FOR
  for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k) {               // :958-981
    int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
    if (rk_step == 0)
      tend_u_euler(e, k) = -cqu(e, k) * ((pressure_p(cell2, k) - pressure_p(cell1, k)) * o->invDcEdge[e]
                            / (0.5 * (zz(cell2, k) + zz(cell1, k)))
                            - 0.5 * zxu(e, k) * (dpdz(cell1, k) + dpdz(cell2, k)));
    wduz(e, k) = 0.0;
    if (k == 1 || k == L - 1)
      wduz(e, k) = 0.5 * (rw(cell1, k) + rw(cell2, k)) * (fzm[k] * u(e, k) + fzp[k] * u(e, k - 1));
    if (k > 1 && k < L - 1)
      wduz(e, k) = flux3(u(e, k - 2), u(e, k - 1), u(e, k), u(e, k + 1), 0.5 * (rw(cell1, k) + rw(cell2, k)), 1.0);
  }
  OMP_FOR
  for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k) {               // :983-1019
    int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
    tend_u(e, k) = -rdzw[k] * (wduz(e, k + 1) - wduz(e, k));
    q(e, k) = 0.0;
    for (int j = 0; j < o->nEdgesOnEdge[e]; ++j) {
      int eoe = o->edgesOnEdge[e * ME2 + j];
      for (int kk = 0; kk < L; ++kk) {                                     // Q14: L identical terms
        double workpv = 0.5 * (pv_edge(e, k) + pv_edge(eoe, k));
        q(e, k) += o->weightsOnEdge[e * ME2 + j] * u(eoe, k) * workpv;
      }
    }
    tend_u(e, k) += rho_edge(e, k) * (q(e, k) - (ke(cell2, k) - ke(cell1, k)) * o->invDcEdge[e])
                    - u(e, k) * 0.5 * (h_divergence(cell1, k) + h_divergence(cell2, k));
    tend_u(e, k) -= (2.0 * C.omega * cos(o->angleEdge[e]) * cos(o->latEdge[e]) * rho_edge(e, k)
                     * 0.25 * (w(cell1, k) + w(cell1, k + 1) + w(cell2, k) + w(cell2, k + 1)))
                    - (u(e, k) * 0.25 * (w(cell1, k) + w(cell1, k + 1) + w(cell2, k) + w(cell2, k + 1)) * rho_edge(e, k) * inv_r_earth);
  }

  if (rk_step == 0) {                                                       // :1025
    OMP_FOR
    for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k) {             // :1030-1048
      delsq_u(e, k) = 0.0;
      int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
      int vertex1 = o->verticesOnEdge[e * 2 + 0], vertex2 = o->verticesOnEdge[e * 2 + 1];
      double r_dc = o->invDcEdge[e];
      double r_dv = std::min(o->invDvEdge[e], 4 * r_dc);
      double u_diffusion = (divergence(cell2, k) - divergence(cell1, k)) * r_dc - (vorticity(vertex2, k) - vorticity(vertex1, k)) * r_dv;
      delsq_u(e, k) += u_diffusion;
      double kdiffu = 0.5 * (kdiff(cell1, k) + kdiff(cell2, k));
      tend_u_euler(e, k) += rho_edge(e, k) * kdiffu * u_diffusion * o->meshScalingDel2[e];
    }
    if (h_mom_eddy_visc4 > 0.0) {                                           // :1050-1091
      OMP_FOR
      for (int iv = 0; iv < nV; ++iv) for (int k = 0; k < L; ++k) {
        delsq_vorticity(iv, k) = 0.0;
        for (int i = 0; i < VD; ++i) {
          int e = o->edgesOnVertex[iv * VD + i];
          double edge_sign = o->invAreaTriangle[iv] * o->dcEdge[e] * o->edgesOnVertex_sign[iv * VD + i];
          delsq_vorticity(iv, k) += edge_sign * delsq_u(e, k);
        }
      }
      OMP_FOR
      for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {
        delsq_divergence(c, k) = 0.0;
        double r = o->invAreaCell[c];
        for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
          int e = o->edgesOnCell[c * ME + i];
          double edge_sign = r * o->dvEdge[e] * o->edgesOnCell_sign[c * ME + i];
          delsq_divergence(c, k) += edge_sign * delsq_u(e, k);
        }
      }
      OMP_FOR
      for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k) {
        int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
        int vertex1 = o->verticesOnEdge[e * 2 + 0], vertex2 = o->verticesOnEdge[e * 2 + 1];
        double u_mix_scale = o->meshScalingDel4[e] * h_mom_eddy_visc4;
        double r_dc = u_mix_scale * C.config_del4u_div_factor * o->invDcEdge[e];
        double r_dv = u_mix_scale * std::min(o->invDvEdge[e], 4 * o->invDcEdge[e]);
        double u_diffusion = rho_edge(e, k) * ((delsq_divergence(cell2, k) - delsq_divergence(cell1, k)) * r_dc
                                               - (delsq_vorticity(vertex2, k) - delsq_vorticity(vertex1, k)) * r_dv);
        tend_u_euler(e, k) -= u_diffusion;
      }
    }
    if (v_mom_eddy_visc2 > 0.0) {                                           // :1094-1146
      if (config_mix_full) {
        OMP_FOR
        for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k) {
          int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
          if (k > 0 && k < L - 1) {
            double z1 = 0.5 * (zgrid(cell1, k - 1) + zgrid(cell2, k - 1)), z2 = 0.5 * (zgrid(cell1, k) + zgrid(cell2, k));
            double z3 = 0.5 * (zgrid(cell1, k + 1) + zgrid(cell2, k + 1)), z4 = 0.5 * (zgrid(cell1, k + 2) + zgrid(cell2, k + 2));
            double zm = 0.5 * (z1 + z2), z0 = 0.5 * (z2 + z3), zp = 0.5 * (z3 + z4);
            tend_u_euler(e, k) += rho_edge(e, k) * v_mom_eddy_visc2 * ((u(e, k + 1) - u(e, k)) / (zp - z0) - (u(e, k) - u(e, k - 1)) / (z0 - zm)) / (0.5 * (zp - zm));
          }
        }
      } else {
        OMP_FOR
        for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k)
          u_mix(e, k) = u(e, k) - u_init[k] * cos(o->angleEdge[e]) - v_init[k] * sin(o->angleEdge[e]);
        OMP_FOR
        for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k) {
          int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
          if (k > 0 && k < L - 1) {
            double z1 = 0.5 * (zgrid(cell1, k - 1) + zgrid(cell2, k - 1)), z2 = 0.5 * (zgrid(cell1, k) + zgrid(cell2, k));
            double z3 = 0.5 * (zgrid(cell1, k + 1) + zgrid(cell2, k + 1)), z4 = 0.5 * (zgrid(cell1, k + 2) + zgrid(cell2, k + 2));
            double zm = 0.5 * (z1 + z2), z0 = 0.5 * (z2 + z3), zp = 0.5 * (z3 + z4);
            tend_u_euler(e, k) += rho_edge(e, k) * v_mom_eddy_visc2 * ((u_mix(e, k + 1) - u_mix(e, k)) / (zp - z0) - (u_mix(e, k) - u_mix(e, k - 1)) / (z0 - zm)) / (0.5 * (zp - zm));
          }
        }
      }
    }
  }

  if (config_rayleigh_damp_u) {                                             // :1152-1159
    double rayleigh_coef_inverse = 1.0 / ((double)(C.config_number_rayleigh_damp_u_levels) * (C.config_rayleigh_damp_u_timescale_days * 86400.0));
    OMP_FOR
    for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k)
      if (k > L - C.config_number_rayleigh_damp_u_levels + 1) {
        double coef = (double)((double)k - (L - C.config_number_rayleigh_damp_u_levels)) * rayleigh_coef_inverse;   // :792-796
        tend_u(e, k) -= rho_edge(e, k) * u(e, k) * coef;
      }
  }
  OMP_FOR
  for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k) tend_u(e, k) += tend_u_euler(e, k) + tend_ru_physics(e, k);   // :1161-1163

  // -------- W section (cr.w is the w tendency here, Q17)
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) w(c, k) = 0.0;   // :1170-1172
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {               // :1174-1197 (Q18: per-point fields overwritten per edge)
    for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
      int e = o->edgesOnCell[c * ME + i];
      double edge_sign = o->edgesOnCell_sign[c * ME + i] * o->dvEdge[e] * 0.5; (void)edge_sign;
      if (k > 0) ru_edge_w(c, k) = fzm[k] * ru(e, k) + fzp[k] * ru(e, k - 1);
      for (int kk = 0; kk < L; ++kk) flux_arr(c, k) = 0.0;
      for (int j = 0; j < o->nAdvCellsForEdge[e]; ++j) {
        int iAdvCell = o->advCellsForEdge[e * NA + j];
        if (k > 0) {
          double scalar_weight = o->adv_coefs[e * NA + j] + copysign(1.0, ru_edge_w(c, k)) * o->adv_coefs_3rd[e * NA + j];
          flux_arr(c, k) += scalar_weight * w(iAdvCell, k);
        }
      }
    }
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k)                 // :1199-1205
    for (int i = 0; i < o->nEdgesOnCell[c]; ++i)
      if (k > 0) w(c, k) -= o->edgesOnCell_sign[c * ME + i] * ru_edge_w(c, k) * flux_arr(c, k);
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) if (k > 0) {    // :1208-1218
    w(c, k) += (rho_zz(c, k) * fzm[k] + rho_zz(c, k - 1) * fzp[k])
               * (pow(fzm[k] * uReconstructZonal(c, k) + fzp[k] * uReconstructZonal(c, k - 1), 2.0)
                  + pow(fzm[k] * uReconstructMeridional(c, k) + fzp[k] * uReconstructMeridional(c, k - 1), 2.0)) / r_earth
               + 2.0 * C.omega * cos(o->latCell[c])
               * (fzm[k] * uReconstructZonal(c, k) + fzp[k] * uReconstructZonal(c, k - 1))
               * (rho_zz(c, k) * fzm[k] + rho_zz(c, k - 1) * fzp[k]);
  }
  if (rk_step == 0) {                                                       // :1224-1274
    OMP_FOR
    for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {
      delsq_w(c, k) = 0.0;
      tend_w_euler(c, k) = 0.0;
      double r_areaCell = o->invAreaCell[c];
      for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
        int e = o->edgesOnCell[c * ME + i];
        double edge_sign = 0.5 * r_areaCell * o->edgesOnCell_sign[c * ME + i] * o->dvEdge[e] * o->invDcEdge[e];
        int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
        if (k > 0) {
          double w_turb_flux = edge_sign * (rho_edge(e, k) + rho_edge(e, k - 1)) * (w(cell2, k) - w(cell1, k));
          delsq_w(c, k) += w_turb_flux;
          w_turb_flux *= o->meshScalingDel2[e] * 0.25 * (kdiff(cell1, k) + kdiff(cell2, k) + kdiff(cell1, k - 1) + kdiff(cell2, k - 1));
          tend_w_euler(c, k) += w_turb_flux;
        }
      }
    }
    if (h_mom_eddy_visc4 > 0.0) {
      OMP_FOR
      for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {
        double r_areaCell = h_mom_eddy_visc4 * o->invAreaCell[c];
        for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
          int e = o->edgesOnCell[c * ME + i];
          int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
          double edge_sign = o->meshScalingDel4[e] * r_areaCell * o->dvEdge[e] * o->edgesOnCell_sign[c * ME + i] * o->invDcEdge[e];
          if (k > 0) tend_w_euler(c, k) -= edge_sign * (delsq_w(cell2, k) - delsq_w(cell1, k));
        }
      }
    }
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {               // :1277-1287
    wdwz(c, k) = 0.0;
    if (k == 1 || k == L - 1) wdwz(c, k) = 0.25 * (rw(c, k) + rw(c, k - 1)) * (w(c, k) + w(c, k - 1));
    if (k > 1 && k < L - 1) wdwz(c, k) = flux3(w(c, k - 2), w(c, k - 1), w(c, k), w(c, k + 1), 0.5 * (rw(c, k) + rw(c, k - 1)), 1.0);
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {               // :1289-1302 (Q19: x *= a - b)
    if (k > 0) w(c, k) *= o->invAreaCell[c] - rdzu[k] * (wdwz(c, k + 1) - wdwz(c, k));
    if (rk_step == 0 && k > 0)
      tend_w_euler(c, k) -= cqw(c, k) * (rdzu[k] * (pressure_p(c, k) - pressure_p(c, k - 1)) - (fzm[k] * dpdz(c, k) + fzp[k] * dpdz(c, k - 1)));
  }
  if (rk_step == 0 && v_mom_eddy_visc2 > 0.0) {                            // :1304-1315
    OMP_FOR
    for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) if (k > 0)
      tend_w_euler(c, k) += v_mom_eddy_visc2 * (rho_zz(c, k) + rho_zz(c, k - 1)) * 0.5
                            * ((w(c, k + 1) - w(c, k)) * rdzw[k] - (w(c, k) - w(c, k - 1)) * rdzw[k - 1]) * rdzu[k];
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) if (k > 0) w(c, k) += tend_w_euler(c, k);   // :1318-1322

  // -------- THETA section
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {               // :1328-1344
    tend_theta(c, k) = 0.0;
    for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
      int e = o->edgesOnCell[c * ME + i];
      flux_arr(c, k) = 0.0;
      for (int j = 0; j < o->nAdvCellsForEdge[e]; ++j) {
        int iAdvCell = o->advCellsForEdge[e * NA + j];
        double scalar_weight = o->adv_coefs[e * NA + j] + copysign(1.0, ru(e, k)) * o->adv_coefs_3rd[e * NA + j];
        flux_arr(c, k) += scalar_weight * theta_m(iAdvCell, k);
      }
      tend_theta(c, k) -= o->edgesOnCell_sign[c * ME + i] * ru(e, k) * flux_arr(c, k);
    }
  }
  if (rk_step > 0) {                                                        // :1347-1360
    OMP_FOR
    for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k)
      for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
        int e = o->edgesOnCell[c * ME + i];
        int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
        double flux = o->edgesOnCell_sign[c * ME + i] * o->dvEdge[e] * (ru_save(e, k) - ru(e, k)) * 0.5 * (theta_m_save(cell2, k) + theta_m_save(cell1, k));
        tend_theta(c, k) -= flux;
      }
  }
  if (rk_step == 0) {                                                       // :1364-1401
    OMP_FOR
    for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {
      delsq_theta(c, k) = 0.0;
      tend_theta_euler(c, k) = 0.0;
      double r_areaCell = o->invAreaCell[c];
      for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
        int e = o->edgesOnCell[c * ME + i];
        double edge_sign = r_areaCell * o->edgesOnCell_sign[c * ME + i] * o->dvEdge[e] * o->invDcEdge[e];
        double pr_scale = prandtl_inv * o->meshScalingDel2[e];
        int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
        double theta_turb_flux = edge_sign * (theta_m(cell2, k) - theta_m(cell1, k)) * rho_edge(e, k);
        delsq_theta(c, k) += theta_turb_flux;
        theta_turb_flux *= 0.5 * (kdiff(cell1, k) + kdiff(cell2, k)) * pr_scale;
        tend_theta_euler(c, k) += theta_turb_flux;
      }
    }
    if (h_theta_eddy_visc4 > 0.0) {
      OMP_FOR
      for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {
        double r_areaCell = h_theta_eddy_visc4 * prandtl_inv * o->invAreaCell[c];
        for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
          int e = o->edgesOnCell[c * ME + i];
          double edge_sign = o->meshScalingDel4[e] * r_areaCell * o->dvEdge[e] * o->edgesOnCell_sign[c * ME + i] * o->invDcEdge[e];
          int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
          tend_theta_euler(c, k) -= edge_sign * (delsq_theta(cell2, k) - delsq_theta(cell1, k));
        }
      }
    }
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {               // :1406-1420
    wdtz(c, k) = 0.0;
    if (k > 0 && k < L - 1) wdtz(c, k) = ((rw_save(c, k) - rw(c, k)) * (fzm[k] * theta_m_save(c, k) + fzp[k] * theta_m_save(c, k - 1)));
    if (k == 1) wdtz(c, k) += rw(c, k) * (fzm[k] * theta_m(c, k) + fzp[k] * theta_m(c, k - 1));
    if (k == L - 1) wdtz(c, k) = rw_save(c, k) * (fzm[k] * theta_m_save(c, k) + fzp[k] * theta_m_save(c, k - 1));
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {               // :1422-1427 (Q19)
    tend_theta(c, k) *= o->invAreaCell[c] - rdzw[k] * (wdtz(c, k + 1) - wdtz(c, k));
    tend_rtheta_adv(c, k) = tend_theta(c, k);
    rthdynten(c, k) = tend_theta(c, k) / rho_zz(c, k);
    tend_theta(c, k) += rho_zz(c, k) * rt_diabatic_tend(c, k);
  }
  if (rk_step == 0 && v_theta_eddy_visc2 > 0.0) {                          // :1430-1475
    OMP_FOR
    for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) if (k > 0 && k < L - 1) {
      double z1 = zgrid(c, k - 1), z2 = zgrid(c, k), z3 = zgrid(c, k + 1), z4 = zgrid(c, k + 2);
      double zm = 0.5 * (z1 + z2), z0 = 0.5 * (z2 + z3), zp = 0.5 * (z3 + z4);
      if (config_mix_full)
        tend_theta_euler(c, k) += v_theta_eddy_visc2 * prandtl_inv * rho_zz(c, k)
                                  * ((theta_m(c, k + 1) - theta_m(c, k)) / (zp - z0) - (theta_m(c, k) - theta_m(c, k - 1)) / (z0 - zm)) / (0.5 * (zp - zm));
      else
        tend_theta_euler(c, k) += v_theta_eddy_visc2 * prandtl_inv * rho_zz(c, k)
                                  * (((theta_m(c, k + 1) - t_init(c, k + 1)) - (theta_m(c, k) - t_init(c, k))) / (zp - z0)
                                     - ((theta_m(c, k) - t_init(c, k)) - (theta_m(c, k - 1) - t_init(c, k - 1))) / (z0 - zm)) / (0.5 * (zp - zm));
    }
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) tend_theta(c, k) += tend_theta_euler(c, k) + tend_rtheta_physics(c, k);   // :1477-1479
  return 0;
}

// atm_set_smlstep_pert_variables_work -- dynamics_tasks.rg:1503-1528 (every point of cpr, level 0 included, Q22/Q23)
int oracle_set_smlstep_pert_variables(oracle_t* o) {
  const int L = o->L, nC = o->nCells, ME = o->maxEdges;
  VF(fzm); VF(fzp);
  CF(w); CF(zz); CF(u_tend);
  F3A zb_cell = o->fa(MPASB200_F_zb_cell), zb3_cell = o->fa(MPASB200_F_zb3_cell);
  OMP_FOR
  for (int c = 0; c < nC; ++c) {
    if (!o->inCpr[c]) continue;
    // `for iCell in cpr` (:1516) visits every point of the region, i.e. levels 0..nVertLevels INCLUSIVE (the index
    // spaces have nVertLevels+1 levels, main.rg:21-24): level nVertLevels is processed with whatever fzm/fzp/zz/
    // zb_cell/u_tend hold there (zero until written, rule M1 -- then w(:, nVertLevels) becomes 0 every stage).
    for (int k = 0; k <= L; ++k) {
      if (o->bdyMaskCell[c] <= o->c.nRelaxZone) {
        for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
          int e = o->edgesOnCell[c * ME + i];
          double flux = o->edgesOnCell_sign[c * ME + i] * (fzm[k] * u_tend(e, k) + fzp[k] * u_tend(e, k - 1));
          w(c, k) -= (zb_cell(c, k, i) + copysign(1.0, u_tend(e, k)) * zb3_cell(c, k, i)) * flux;
        }
        w(c, k) *= (fzm[k] * zz(c, k) + fzp[k] * zz(c, k - 1));
      }
    }
  }
  return 0;
}

// MPASB200_PHYSICS_CORRECTED form of atm_advance_acoustic_step_work (mpas_b200.h).  The three
// pieces the reference has commented out are enabled with the semantics of MPAS v7's routine of the same name:
//   * edge loop  :1585-1613  exactly the commented lines (ru_p, ruAvg), every edge (the nCellsSolve test stays out);
//   * cell loop  :1632-1704  column by column: rs/ts live for the whole column (the reference zeroes them at every
//     point only because its loop variable is a point, Q25); forward elimination :1670-1671 for all levels, then the
//     back-substitution :1674-1677 `rw_p(k) -= gamma_tri(k) * rw_p(k+1)` for k = nVertLevels-1 .. 0 (Fortran
//     k = nVertLevels,1,-1), then Rayleigh damping :1681-1690, then rho_pp / rtheta_pp :1694-1696.
// Field bindings stay the reference's (cr.w for tend_rw, cr.theta_m for tend_rt, er.tend_ru).
static int acoustic_step_corrected(oracle_t* o, double dts, int small_step) {
  const int L = o->L, nC = o->nCells, nE = o->nEdges, ME = o->maxEdges;
  const double epssm = o->c.config_epssm, rgas = o->c.rgas, gravity = o->c.gravity;
  const double rcv = rgas / (o->c.cp - rgas); const double c2 = o->c.cp * rcv;
  const double resm = (1.0 - epssm) / (1.0 + epssm);
  VF(cofrz); VF(fzm); VF(fzp); VF(rdzw);
  CF(a_tri); CF(alpha_tri); CF(gamma_tri); CF(coftz); CF(cofwr); CF(cofwt); CF(cofwz); CF(dss); CF(rho_pp); CF(rho_zz); CF(rtheta_pp);
  CF(rw); CF(rw_save); CF(tend_rho); CF(theta_m); CF(w); CF(zz); CF(rtheta_pp_old); CF(rw_p); CF(wwAvg); CF(ru_p);
  CF(exner); CF(cqu); CF(zxu); CF(tend_ru); CF(ruAvg);
  if (small_step != 0) {                                                    // :1581-1599
    OMP_FOR
    for (int e = 0; e < nE; ++e) {
      if (!o->edge_on(e)) continue;
      int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
      for (int k = 0; k < L; ++k) {
        double pgrad = ((rtheta_pp(cell2, k) - rtheta_pp(cell1, k)) * o->invDcEdge[e]) / (0.5 * (zz(cell2, k) + zz(cell1, k)));   // :1591
        pgrad *= cqu(e, k) * 0.5 * c2 * (exner(cell1, k) + exner(cell2, k));                                                  // :1592
        pgrad += 0.5 * zxu(e, k) * gravity * (rho_pp(cell1, k) + rho_pp(cell2, k));                                           // :1593
        ru_p(e, k) += dts * (tend_ru(e, k) - (1.0 - o->specZoneMaskEdge[e]) * pgrad);                                          // :1594
        ruAvg(e, k) += ru_p(e, k);                                                                                             // :1597
      }
    }
  } else {                                                                  // :1601-1613
    OMP_FOR
    for (int e = 0; e < nE; ++e) if (o->edge_on(e)) for (int k = 0; k < L; ++k) { ru_p(e, k) = dts * tend_ru(e, k); ruAvg(e, k) = ru_p(e, k); }
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) {
    if (!o->cell_on(c)) continue;
    std::vector<double> rs(L), ts(L);
    for (int k = 0; k < L; ++k) rtheta_pp_old(c, k) = (small_step == 0) ? 0.0 : rtheta_pp(c, k);       // :1615-1623
    if (small_step == 0) {                                                  // :1625-1636
      for (int k = 0; k <= L; ++k) { wwAvg(c, k) = 0; rw_p(c, k) = 0; }
      for (int k = 0; k < L; ++k) { rho_pp(c, k) = 0; rtheta_pp(c, k) = 0; }
    }
    if (o->specZoneMaskCell[c] == 0.0) {                                    // :1638
      for (int k = 0; k < L; ++k) { ts[k] = 0; rs[k] = 0; }
      for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {                        // :1644-1652
        int e = o->edgesOnCell[c * ME + i];
        int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
        for (int k = 0; k < L; ++k) {
          double flux = o->edgesOnCellSign[c * ME + i] * dts * o->dvEdge[e] * ru_p(e, k) * o->invAreaCell[c];
          rs[k] -= flux;
          ts[k] -= flux * 0.5 * (theta_m(cell2, k) + theta_m(cell1, k));
        }
      }
      for (int k = 0; k < L; ++k) {                                         // :1657-1658
        rs[k] = rho_pp(c, k) + dts * tend_rho(c, k) + rs[k] - cofrz[k] * resm * (rw_p(c, k + 1) - rw_p(c, k));
        ts[k] = rtheta_pp(c, k) + dts * theta_m(c, k) + ts[k] - resm * rdzw[k] * (coftz(c, k + 1) * rw_p(c, k + 1) - coftz(c, k) * rw_p(c, k));
      }
      for (int k = 1; k < L; ++k) {                                         // :1660-1667
        wwAvg(c, k) += 0.5 * (1.0 - epssm) * rw_p(c, k);
        rw_p(c, k) += dts * w(c, k) - cofwz(c, k)
                      * ((zz(c, k) * ts[k] - zz(c, k - 1) * ts[k - 1]) + resm * (zz(c, k) * rtheta_pp(c, k) - zz(c, k - 1) * rtheta_pp(c, k - 1)))
                      - cofwr(c, k) * ((rs[k] + rs[k - 1]) + resm * (rho_pp(c, k) + rho_pp(c, k - 1)))
                      + cofwt(c, k) * (ts[k] + resm * rtheta_pp(c, k))
                      + cofwt(c, k - 1) * (ts[k - 1] + resm * rtheta_pp(c, k - 1));
      }
      for (int k = 1; k < L; ++k) {                                         // :1670-1671
        rw_p(c, k) -= a_tri(c, k) * rw_p(c, k - 1);
        rw_p(c, k) *= alpha_tri(c, k);
      }
      for (int k = L - 1; k >= 0; --k) rw_p(c, k) -= gamma_tri(c, k) * rw_p(c, k + 1);                 // :1674-1677
      for (int k = 1; k < L; ++k) {                                         // :1681-1690
        rw_p(c, k) += (rw_save(c, k) - rw(c, k)) - dts * dss(c, k)
                      * (fzm[k] * zz(c, k) + fzp[k] * zz(c, k - 1)) * (fzm[k] * rho_zz(c, k) + fzp[k] * rho_zz(c, k - 1)) * w(c, k);
        rw_p(c, k) /= (1.0 + dts * dss(c, k));
        rw_p(c, k) -= (rw_save(c, k) - rw(c, k));
        wwAvg(c, k) += 0.5 * (1.0 + epssm) * rw_p(c, k);
      }
      for (int k = 0; k < L; ++k) {                                         // :1694-1696
        rho_pp(c, k) = rs[k] - cofrz[k] * (rw_p(c, k + 1) - rw_p(c, k));
        rtheta_pp(c, k) = ts[k] - rdzw[k] * (coftz(c, k + 1) * rw_p(c, k + 1) - coftz(c, k) * rw_p(c, k));
      }
    } else {                                                                // :1698-1703
      for (int k = 0; k < L; ++k) {
        rho_pp(c, k) = rho_pp(c, k) + dts * tend_rho(c, k);
        rtheta_pp(c, k) = rtheta_pp(c, k) + dts * theta_m(c, k);
        rw_p(c, k) = rw_p(c, k) + dts * w(c, k);
        wwAvg(c, k) = wwAvg(c, k) + 0.5 * (1.0 + epssm) * rw_p(c, k);
      }
    }
  }
  return 0;
}

// atm_advance_acoustic_step_work -- dynamics_tasks.rg:1546-1705
int oracle_advance_acoustic_step(oracle_t* o, double dts, int small_step) {
  if (o->c.physics_mode == MPASB200_PHYSICS_CORRECTED) return acoustic_step_corrected(o, dts, small_step);
  const int L = o->L, nC = o->nCells, ME = o->maxEdges;
  const double epssm = o->c.config_epssm, rgas = o->c.rgas;
  const double rcv = rgas / (o->c.cp - rgas); const double c2 = o->c.cp * rcv; (void)c2;
  const double resm = (1.0 - epssm) / (1.0 + epssm);
  VF(cofrz); VF(fzm); VF(fzp); VF(rdzw);
  CF(a_tri); CF(alpha_tri); CF(coftz); CF(cofwr); CF(cofwt); CF(cofwz); CF(dss); CF(rho_pp); CF(rho_zz); CF(rtheta_pp);
  CF(rw); CF(rw_save); CF(tend_rho); CF(theta_m); CF(w); CF(zz); CF(rtheta_pp_old); CF(rw_p); CF(wwAvg); CF(ru_p);
  // the u / ru_p update is commented out in the reference (:1585-1613, Q24): both edge loops are empty.
  if (small_step == 0) {                                                    // :1615-1623
    OMP_FOR
    for (int c = 0; c < nC; ++c) if (o->cell_on(c)) for (int k = 0; k < L; ++k) rtheta_pp_old(c, k) = 0;
  } else {
    OMP_FOR
    for (int c = 0; c < nC; ++c) if (o->cell_on(c)) for (int k = 0; k < L; ++k) rtheta_pp_old(c, k) = rtheta_pp(c, k);
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) if (o->cell_on(c)) for (int k = 0; k <= L; ++k) if (small_step == 0) { wwAvg(c, k) = 0; rw_p(c, k) = 0; }   // :1625-1630
  OMP_FOR
  for (int c = 0; c < nC; ++c) {
    if (!o->cell_on(c)) continue;
    std::vector<double> rs(L), ts(L);                                       // task-level scratch, re-zeroed at every point (Q25)
    for (int k = 0; k < L; ++k) {                                           // :1632-1704, levels ascending (M4)
      if (small_step == 0) { rho_pp(c, k) = 0; rtheta_pp(c, k) = 0; }
      if (o->specZoneMaskCell[c] == 0.0) {                                  // :1638 (Q26)
        for (int i = 0; i < L; ++i) { ts[i] = 0; rs[i] = 0; }
        for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {                      // :1644-1652
          int e = o->edgesOnCell[c * ME + i];
          int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
          double flux = o->edgesOnCellSign[c * ME + i] * dts * o->dvEdge[e] * ru_p(e, k) * o->invAreaCell[c];
          rs[k] -= flux;
          ts[k] -= flux * 0.5 * (theta_m(cell2, k) + theta_m(cell1, k));
        }
        rs[k] = rho_pp(c, k) + dts * tend_rho(c, k) + rs[k] - cofrz[k] * resm * (rw_p(c, k + 1) - rw_p(c, k));               // :1657
        ts[k] = rtheta_pp(c, k) + dts * theta_m(c, k) + ts[k] - resm * rdzw[k] * (coftz(c, k + 1) * rw_p(c, k + 1) - coftz(c, k) * rw_p(c, k));   // :1658 (theta_m for tend_rt, Q27)
        if (k > 0) {                                                        // :1660-1672
          wwAvg(c, k) += 0.5 * (1.0 - epssm) * rw_p(c, k);
          rw_p(c, k) += dts * w(c, k) - cofwz(c, k)
                        * ((zz(c, k) * ts[k] - zz(c, k - 1) * ts[k - 1]) + resm * (zz(c, k) * rtheta_pp(c, k) - zz(c, k - 1) * rtheta_pp(c, k - 1)))
                        - cofwr(c, k) * ((rs[k] + rs[k - 1]) + resm * (rho_pp(c, k) + rho_pp(c, k - 1)))
                        + cofwt(c, k) * (ts[k] + resm * rtheta_pp(c, k))
                        + cofwt(c, k - 1) * (ts[k - 1] + resm * rtheta_pp(c, k - 1));
          rw_p(c, k) -= a_tri(c, k) * rw_p(c, k - 1);
          rw_p(c, k) *= alpha_tri(c, k);
        }
        // back-substitution is commented out in the reference (:1674-1677, Q28)
        if (k > 0) {                                                        // :1681-1690
          rw_p(c, k) += (rw_save(c, k) - rw(c, k)) - dts * dss(c, k)
                        * (fzm[k] * zz(c, k) + fzp[k] * zz(c, k - 1)) * (fzm[k] * rho_zz(c, k) + fzp[k] * rho_zz(c, k - 1)) * w(c, k);
          rw_p(c, k) /= (1.0 + dts * dss(c, k));
          rw_p(c, k) -= (rw_save(c, k) - rw(c, k));
          wwAvg(c, k) += 0.5 * (1.0 + epssm) * rw_p(c, k);
        }
        rho_pp(c, k) = rs[k] - cofrz[k] * (rw_p(c, k + 1) - rw_p(c, k));   // :1694
        rtheta_pp(c, k) = ts[k] - rdzw[k] * (coftz(c, k + 1) * rw_p(c, k + 1) - coftz(c, k) * rw_p(c, k));   // :1695-1696
      } else {                                                              // :1698-1703
        rho_pp(c, k) = rho_pp(c, k) + dts * tend_rho(c, k);
        rtheta_pp(c, k) = rtheta_pp(c, k) + dts * theta_m(c, k);
        rw_p(c, k) = rw_p(c, k) + dts * w(c, k);
        wwAvg(c, k) = wwAvg(c, k) + 0.5 * (1.0 + epssm) * rw_p(c, k);
      }
    }
  }
  return 0;
}

// atm_divergence_damping_3d -- dynamics_tasks.rg:1726-1763
int oracle_divergence_damping_3d(oracle_t* o, double dts) {
  const int L = o->L, nE = o->nEdges;
  CF(rtheta_pp); CF(rtheta_pp_old); CF(theta_m); CF(ru_p);
  double smdiv = o->c.config_smdiv;
  double rdts = 1.0 / dts;
  double coef_divdamp = 2.0 * smdiv * o->c.config_len_disp * rdts;
  OMP_FOR
  for (int e = 0; e < nE; ++e) {
    if (!o->edge_on(e)) continue;
    int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
    if (!(o->isShared[cell1] && o->isShared[cell2])) {
      for (int k = 0; k < L; ++k) {
        double divCell1 = -(rtheta_pp(cell1, k) - rtheta_pp_old(cell1, k));
        double divCell2 = -(rtheta_pp(cell2, k) - rtheta_pp_old(cell2, k));
        ru_p(e, k) += coef_divdamp * (divCell2 - divCell1) * (1.0 - o->specZoneMaskEdge[e]) / (theta_m(cell1, k) + theta_m(cell2, k));
      }
    }
  }
  return 0;
}

// atm_recover_large_step_variables_work -- dynamics_tasks.rg:1766-1872 (not called by atm_srk3: rk_timestep.rg:460)
int oracle_recover_large_step_variables(oracle_t* o, int ns, int rk_step, double dt) {
  const int L = o->L, nC = o->nCells, nE = o->nEdges, ME = o->maxEdges;
  const double rgas = o->c.rgas, rcv = rgas / (o->c.cp - rgas);
  const int p0 = 100000;
  VF(cf1); VF(cf2); VF(cf3); VF(fzm); VF(fzp);
  CF(exner_base); CF(rho_base); CF(rho_p_save); CF(rho_pp); CF(rt_diabatic_tend); CF(rtheta_base); CF(rtheta_p_save); CF(rtheta_pp);
  CF(rw_p); CF(rw_save); CF(zz); CF(ru_p); CF(ru_save); CF(pressure_p); CF(theta_m); CF(u); CF(exner); CF(rho_p); CF(rho_zz);
  CF(rtheta_p); CF(rw); CF(w); CF(wwAvg); CF(ruAvg); CF(ru);
  F3A zb_cell = o->fa(MPASB200_F_zb_cell), zb3_cell = o->fa(MPASB200_F_zb3_cell);
  const bool fix = o->c.physics_mode == MPASB200_PHYSICS_CORRECTED;         // four expressions restored, see mpas_b200.h
  for (int k = 0; k < L; ++k) rho_zz(nC, k) = 1.0;                          // :1792-1794 the "garbage cell" = the pad cell
  double invNs = 1 / (double)(ns);
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {               // :1800-1826
    rho_p(c, k) = rho_p_save(c, k) + rho_pp(c, k);
    rho_zz(c, k) = rho_p(c, k) + rho_base(c, k);
    w(c, k) = 0.0;
    wwAvg(c, k) *= invNs;
    wwAvg(c, k) += rw_save(c, k);
    rw(c, k) = rw_save(c, k) + rw_p(c, k);
    w(c, k) = (fix && k == 0) ? 0.0 : rw(c, k) / (fzm[k] * zz(c, k) + fzp[k] * zz(c, k - 1));   // :1810; MPAS: w(1) = 0, k = 2..nVertLevels
    if (k == L) w(c, k) = 0.0;                                              // never fires
    if (rk_step == 2) {
      rtheta_p(c, k) = rtheta_p_save(c, k) + rtheta_pp(c, k) - dt * rho_zz(c, k) * rt_diabatic_tend(c, k);
      theta_m(c, k) = (rtheta_p(c, k) + rtheta_base(c, k)) / rho_zz(c, k);
      exner(c, k) = fix ? pow(zz(c, k) * (rgas / p0) * (rtheta_p(c, k) + rtheta_base(c, k)), rcv)
                        : zz(c, k) * (rgas / p0) * pow((rtheta_p(c, k) + rtheta_base(c, k)), rcv);      // :1819
      pressure_p(c, k) = zz(c, k) * rgas * (exner(c, k) * rtheta_p(c, k) + rtheta_base(c, k) * (exner(c, k) - exner_base(c, k)));
    } else {
      rtheta_p(c, k) = rtheta_p_save(c, k) + rtheta_pp(c, k);
      theta_m(c, k) = (rtheta_p(c, k) + rtheta_base(c, k)) / rho_zz(c, k);
    }
  }
  OMP_FOR
  for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k) {               // :1835-1842
    int cell1 = o->cellsOnEdge[e * 2 + 0], cell2 = o->cellsOnEdge[e * 2 + 1];
    ruAvg(e, k) *= invNs;
    ruAvg(e, k) += ru_save(e, k);
    ru(e, k) = fix ? ru_save(e, k) + ru_p(e, k) : ru_save(e, k) * ru_p(e, k);                          // :1840
    u(e, k) = 2 * ru(e, k) / (rho_zz(cell1, k) + rho_zz(cell2, k));
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {               // :1844-1859
    if (o->bdyMaskCell[c] <= o->c.nRelaxZone) {
      for (int i = 0; i < o->nEdgesOnCell[c]; ++i) {
        int e = o->edgesOnCell[c * ME + i];
        double flux = (cf1[0] * ru(e, 0) + cf2[0] * ru(e, 1) + cf3[0] * ru(e, 2));
        w(c, 0) += o->edgesOnCell_sign[c * ME + i] * (zb_cell(c, 0, i) + copysign(1.0, flux) * zb3_cell(c, 0, i)) * flux;
        double flux2 = fix ? fzm[k] * ru(e, k) + fzp[k] * ru(e, k - 1) : fzm[k] * ru(e, k) * (fzp[k] * ru(e, k - 1));   // :1856
        w(c, k) += o->edgesOnCell_sign[c * ME + i] * (zb_cell(c, k, i) + copysign(1.0, flux2) * zb3_cell(c, k, i)) * flux2;
      }
    }
  }
  OMP_FOR
  for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {               // :1861-1871
    if (o->bdyMaskCell[c] <= o->c.nRelaxZone) {
      if (k == 0) w(c, 0) /= (cf1[0] * rho_zz(c, 0) + cf2[0] * rho_zz(c, 1) + cf3[0] * rho_zz(c, 2));
      if (k > 0) w(c, k) /= (fzm[k] * rho_zz(c, k) + fzp[k] * rho_zz(c, k - 1));
    }
  }
  return 0;
}

// atm_rk_dynamics_substep_finish -- dynamics_tasks.rg:1951-2007
int oracle_rk_dynamics_substep_finish(oracle_t* o, int dynamics_substep, int dynamics_split) {
  const int L = o->L, nC = o->nCells, nE = o->nEdges;
  CF(ru); CF(ru_save); CF(u); CF(u_2); CF(rw); CF(rw_save); CF(rtheta_p); CF(rtheta_p_save); CF(rho_p); CF(rho_p_save);
  CF(w); CF(w_2); CF(theta_m); CF(theta_m_2); CF(rho_zz); CF(rho_zz_2); CF(rho_zz_old_split);
  CF(ruAvg); CF(ruAvg_split); CF(wwAvg); CF(wwAvg_split);
  double inv_dynamics_split = 1.0 / (double)(dynamics_split);
  if (dynamics_substep < dynamics_split) {
    OMP_FOR
    for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k) { ru_save(e, k) = ru(e, k); u(e, k) = u_2(e, k); }
    OMP_FOR
    for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) {
      rw_save(c, k) = rw(c, k); rtheta_p_save(c, k) = rtheta_p(c, k); rho_p_save(c, k) = rho_p(c, k);
      w(c, k) = w_2(c, k); theta_m(c, k) = theta_m_2(c, k); rho_zz(c, k) = rho_zz_2(c, k);
    }
  }
  if (dynamics_substep == 1) {
    OMP_FOR
    for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k) ruAvg_split(e, k) = ruAvg(e, k);
    OMP_FOR
    for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) wwAvg_split(c, k) = wwAvg(c, k);
  } else {
    OMP_FOR
    for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k) ruAvg_split(e, k) = ruAvg(e, k) + ruAvg_split(e, k);
    OMP_FOR
    for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) wwAvg_split(c, k) = wwAvg(c, k) + wwAvg_split(c, k);
  }
  if (dynamics_substep == dynamics_split) {
    OMP_FOR
    for (int e = 0; e < nE; ++e) for (int k = 0; k < L; ++k) ruAvg(e, k) = ruAvg_split(e, k) * inv_dynamics_split;
    OMP_FOR
    for (int c = 0; c < nC; ++c) for (int k = 0; k < L; ++k) { wwAvg(c, k) = wwAvg_split(c, k) * inv_dynamics_split; rho_zz(c, k) = rho_zz_old_split(c, k); }
  }
  return 0;
}

// atm_srk3 -- rk_timestep.rg:361-500 (and atm_timestep :503-519)
int oracle_srk3(oracle_t* o, double dt) {
  const MpasConfig& C = o->c;
  int number_of_sub_steps = C.number_of_sub_steps;                          // :378
  int dynamics_split = C.config_dynamics_split_steps;                       // :381
  double dt_dynamics = dt;
  double rk_timestep[3] = {dt_dynamics / 3, dt_dynamics / 2, dt_dynamics};          // :386-389
  double rk_sub_timestep[3] = {dt_dynamics / 3, dt_dynamics / number_of_sub_steps, dt_dynamics / number_of_sub_steps};
  int number_sub_steps[3] = {std::max(1, number_of_sub_steps / 2), std::max(1, number_of_sub_steps / 2), number_of_sub_steps};
  oracle_rk_integration_setup(o);                                           // :404
  oracle_compute_moist_coefficients(o);                                     // :408
  oracle_compute_vert_imp_coefs(o, rk_sub_timestep[0]);                     // :417
  for (int rk_step = 0; rk_step < 3; ++rk_step) {                           // :426
    if (rk_step == 1) oracle_compute_vert_imp_coefs(o, rk_sub_timestep[rk_step]);   // :429-434
    // :437 passes rk_sub_timestep[rk_step] (double) where rk_step:int is expected (Q3)
    int rk_arg = (C.rkarg_policy == MPASB200_RKARG_SUBSTEP_TRUNC) ? (int)rk_sub_timestep[rk_step] : rk_step;
    oracle_compute_dyn_tend(o, rk_arg, dt, C.config_horiz_mixing, C.config_mpas_cam_coef, C.config_mix_full, C.config_rayleigh_damp_u);
    oracle_set_smlstep_pert_variables(o);                                   // :441
    for (int small_step = 0; small_step < number_sub_steps[rk_step] + 1; ++small_step) {   // :450 (Q4)
      oracle_advance_acoustic_step(o, rk_sub_timestep[rk_step], small_step);
      oracle_divergence_damping_3d(o, rk_sub_timestep[rk_step]);
    }
    // :459-460 atm_recover_large_step_variables is commented out (Q5); CORRECTED calls it with the commented arguments
    if (C.physics_mode == MPASB200_PHYSICS_CORRECTED) oracle_recover_large_step_variables(o, number_sub_steps[rk_step], rk_step, dt);
    // :465 "SKIPPING if (config_scalar_advection .and. (.not. config_split_dynamics_transport))": the call MPAS makes here
    if (C.config_scalar_advection) oracle_advance_scalars(o, rk_timestep[rk_step], rk_step);
    oracle_compute_solve_diagnostics(o, 0, rk_step);                        // :467
  }
  oracle_rk_dynamics_substep_finish(o, 1, dynamics_split);                  // :481
  return 0;
}
int oracle_timestep(oracle_t* o, double dt) { return oracle_srk3(o, dt); }

}  // extern "C"
