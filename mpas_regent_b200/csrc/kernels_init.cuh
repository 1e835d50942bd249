// kernels_init.cuh -- the mesh-only producers of atm_core_init on the device:
// atm_compute_signs (dynamics_tasks.rg:46-130), atm_adv_coef_compression (:133-269), atm_couple_coef_3rd_order (:303-325).
// One-time integer / list work over the RAW stored ids in the caller's numbering (they run before the renumbering of
// mpasb200_upload_mesh exists): one thread per vertex / cell / edge, the reference's loops kept as they are -- including the
// quirks (the list index `n`, the exclusive `for j = 0, n` searches, the cap at maxEdges-1).  Index rule M2: `im_R` resolves a
// stored id under the handle's index policy, the pad entity reads as zero.
#pragma once
#include "kernels.cuh"

struct InitMeshDev {
  const int *nEdgesOnCell, *edgesOnCell, *verticesOnCell, *cellsOnCell, *cellsOnEdge, *verticesOnEdge, *cellsOnVertex, *edgesOnVertex;
  const double *dcEdge, *dvEdge, *deriv_two;
  int nC, nE, nV, ME, VD, NA, pol;
};
DI long im_R(const InitMeshDev& M, long id, int n) {
  long idx = (M.pol == MPASB200_INDEX_LITERAL) ? id : id - 1;
  if (idx < 0 || idx > n) idx = n;
  return idx;
}
DI int im_off(const InitMeshDev& M) { return M.pol == MPASB200_INDEX_LITERAL ? 0 : 1; }

// edgesOnVertexSign  :60-72
__global__ void k_signs_vertex(const InitMeshDev M, double* __restrict__ edgesOnVertexSign) {
  const int iVtx = blockIdx.x * blockDim.x + threadIdx.x;
  if (iVtx >= M.nV) return;
  for (int i = 0; i < M.VD; ++i) {
    const int e = M.edgesOnVertex[(size_t)iVtx * M.VD + i];
    double s = 0.0;
    if (e <= M.nE) {
      const long ei = im_R(M, e, M.nE);
      const int v1 = ei < M.nE ? M.verticesOnEdge[ei * 2 + 1] : 0;
      s = (iVtx + im_off(M) == v1) ? 1.0 : -1.0;
    }
    edgesOnVertexSign[(size_t)iVtx * M.VD + i] = s;
  }
}

// edgesOnCellSign  :74-86  and kiteForCell  :113-129 (j runs 1..vertexDegree-1 only; no match leaves the zero)
__global__ void k_signs_cell(const InitMeshDev M, double* __restrict__ edgesOnCellSign, int* __restrict__ kiteForCell) {
  const int iCell = blockIdx.x * blockDim.x + threadIdx.x;
  if (iCell >= M.nC) return;
  const int n = min(M.nEdgesOnCell[iCell], M.ME);
  for (int i = 0; i < M.ME; ++i) {
    double s = 0.0;
    int kite = 0;
    if (i < n) {
      const int e = M.edgesOnCell[(size_t)iCell * M.ME + i];
      if (e <= M.nE) {
        const long ei = im_R(M, e, M.nE);
        const int c1 = ei < M.nE ? M.cellsOnEdge[ei * 2] : 0;
        s = (iCell + im_off(M) == c1) ? 1.0 : -1.0;
      }
      const int iVtx = M.verticesOnCell[(size_t)iCell * M.ME + i];
      if (iVtx <= M.nV) {
        const long vi = im_R(M, iVtx, M.nV);
        for (int j = 1; j < M.VD; ++j) {
          const int c = vi < M.nV ? M.cellsOnVertex[vi * M.VD + j] : 0;
          if (iCell + im_off(M) == c) { kite = j; break; }
        }
      } else kite = 1;
    }
    if (edgesOnCellSign) edgesOnCellSign[(size_t)iCell * M.ME + i] = s;
    if (kiteForCell) kiteForCell[(size_t)iCell * M.ME + i] = kite;
  }
}

// atm_adv_coef_compression  :133-269, one thread per edge.  W = 2 + 2*maxEdges list entries (the reference's cell_list has maxEdges
// and is overrun; the host producer and the oracle use the same wider list), outputs are the first FIFTEEN.
template <int WMAX>
__global__ void k_adv_coef(const InitMeshDev M, int* __restrict__ nAdvCellsForEdge, int* __restrict__ advCellsForEdge,
                           double* __restrict__ adv_coefs, double* __restrict__ adv_coefs_3rd) {
  const int iEdge = blockIdx.x * blockDim.x + threadIdx.x;
  if (iEdge >= M.nE) return;
  const int NA = M.NA, ME = M.ME;
  int cell_list[WMAX];
  double a[WMAX], a3[WMAX];
  for (int j = 0; j < WMAX; ++j) { cell_list[j] = 0; a[j] = 0.0; a3[j] = 0.0; }
  int n = 0;
  const int cell1 = M.cellsOnEdge[(size_t)iEdge * 2], cell2 = M.cellsOnEdge[(size_t)iEdge * 2 + 1];
  const bool on = (cell1 <= M.nC || cell2 <= M.nC);
  if (on) {
    const long i1 = im_R(M, cell1, M.nC), i2 = im_R(M, cell2, M.nC);
    const int n1 = i1 < M.nC ? M.nEdgesOnCell[i1] : 0, n2 = i2 < M.nC ? M.nEdgesOnCell[i2] : 0;
    auto coc = [&](long c, int i) { return c < M.nC ? M.cellsOnCell[c * ME + i] : 0; };
    auto d2 = [&](int idx) { return (M.deriv_two && idx < 2 * NA) ? M.deriv_two[(size_t)iEdge * 2 * NA + idx] : 0.0; };
    cell_list[0] = cell1; cell_list[1] = cell2;
    n = 1;
    for (int i = 0; i < n1; ++i)
      if (coc(i1, i) != cell2) { n += 1; cell_list[n] = coc(i1, i); }
    for (int iCell = 0; iCell < n2; ++iCell) {
      bool addcell = true;
      for (int i = 0; i < n; ++i) if (cell_list[i] == coc(i2, iCell)) addcell = false;
      if (addcell && n < ME - 1) { n += 1; cell_list[n] = coc(i2, iCell); }
    }
    auto j_in_of = [&](int target) { int j_in = 0; for (int j = 0; j < n; ++j) if (cell_list[j] == target) j_in = j; return j_in; };
    int j_in = j_in_of(cell1);
    a[j_in] += d2(0); a3[j_in] += d2(0);
    for (int iCell = 0; iCell < n1; ++iCell) {
      j_in = j_in_of(coc(i1, iCell));
      a[j_in] += d2(iCell * NA + 0); a3[j_in] += d2(iCell * NA + 0);
    }
    j_in = j_in_of(cell2);
    a[j_in] += d2(1); a3[j_in] += d2(1);
    for (int iCell = 0; iCell < n2; ++iCell) {
      j_in = j_in_of(coc(i2, iCell));
      a[j_in] += d2(iCell * NA + 1); a3[j_in] += d2(iCell * NA + 1);
    }
    const double dc = M.dcEdge[iEdge], dv = M.dvEdge[iEdge];
    for (int j = 0; j < n; ++j) {
      a[j] = -1.0 * (dc * dc) * a[j] / 12;          // pow(dcEdge, 2) is exactly dc*dc
      a3[j] = -1.0 * (dc * dc) * a3[j] / 12;
    }
    a[j_in_of(cell1)] += 0.5;
    a[j_in_of(cell2)] += 0.5;
    for (int j = 0; j < n; ++j) { a[j] *= dv; a3[j] *= dv; }
  }
  nAdvCellsForEdge[iEdge] = n;
  for (int j = 0; j < NA; ++j) {
    advCellsForEdge[(size_t)iEdge * NA + j] = (on && j < n && j < WMAX) ? cell_list[j] : 0;
    adv_coefs[(size_t)iEdge * NA + j] = j < WMAX ? a[j] : 0.0;
    adv_coefs_3rd[(size_t)iEdge * NA + j] = j < WMAX ? a3[j] : 0.0;
  }
}

// atm_compute_signs, 3-D part  :88-110 on the device mirror (internal numbering): zb_cell[i] = zb[side] of the cell's slot-i edge,
// side 0 when the cell is that edge's first cell, levels 0..nVertLevels
__global__ void k_zb_cell(const View V) {
  PAIR_THREAD(V.nCells)
  if (!inx || k0 > L) return;
  const bool s1 = k1 <= L;
  const size_t edgeSlot = (size_t)(V.nEdges + 1) * LP;
  const int n = V.nEdgesOnCell[x];
  for (int i = 0; i < n; ++i) {
    const int e = V.edgesOnCell[x * V.MEP + i];
    const size_t side = (x == V.c1OnCell[x * V.MEP + i]) ? 0 : 1;
    st2m(FLD(zb_cell) + (size_t)i * V.cellSlot, ix, ld2(FLD(zb) + side * edgeSlot, (size_t)e * LP + k0), true, s1);
    st2m(FLD(zb3_cell) + (size_t)i * V.cellSlot, ix, ld2(FLD(zb3) + side * edgeSlot, (size_t)e * LP + k0), true, s1);
  }
}

// atm_couple_coef_3rd_order  :303-325
__global__ void k_scale(double* __restrict__ p, size_t n, double c) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] *= c;
}
__global__ void k_zb3_level0(const View V, double c) {          // `cr[{iCell, 0}].zb3_cell[j] *= coef`: level 0 only
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= V.nCells) return;
  for (int j = 0; j < V.maxEdges; ++j) FLD(zb3_cell)[(size_t)j * V.cellSlot + (size_t)x * V.LP] *= c;
}

// atm_compute_mesh_scaling  :595-646 (the del2 / del4 factors; cellOne / cellTwo are cellsOnEdge)
__global__ void k_mesh_scaling(const InitMeshDev M, const double* __restrict__ meshDensity, int scale_with_mesh,
                               double* __restrict__ del2, double* __restrict__ del4) {
  const int iEdge = blockIdx.x * blockDim.x + threadIdx.x;
  if (iEdge >= M.nE) return;
  double a = 1.0, b = 1.0;
  if (scale_with_mesh) {
    const long c1 = im_R(M, M.cellsOnEdge[(size_t)iEdge * 2], M.nC), c2 = im_R(M, M.cellsOnEdge[(size_t)iEdge * 2 + 1], M.nC);
    const double m = ((c1 < M.nC ? meshDensity[c1] : 0.0) + (c2 < M.nC ? meshDensity[c2] : 0.0)) / 2.0;
    a = 1.0 / pow(m, 0.25); b = 1.0 / pow(m, 0.75);
  }
  del2[iEdge] = a; del4[iEdge] = b;
}

// atm_compute_damping_coefs  :274-300 on the device mirror; meshDensity in internal numbering
__global__ void k_damping_coefs(const View V, const double* __restrict__ meshDensity, double config_zd, double config_xnutr) {
  PAIR_THREAD(V.nCells)
  if (!m0) return;
  const double pii = acos(-1.0);
  const double* zg = FLD(zgrid);
  const double zt = zg[(size_t)x * LP + L];
  const double md = pow(meshDensity[x], 0.25);
  D2 out = bc(0.0);
  for (int c = 0; c < 2; ++c) {
    const int k = k0 + c;
    if (k >= L) break;
    const double z = 0.5 * (zg[ix + c] + zg[ix + c + 1]);
    double d = 0.0;
    if (z > config_zd) {
      const double s = sin(0.5 * pii * (z - config_zd) / (zt - config_zd));
      d = config_xnutr * (s * s);          // pow(x, 2.0)
      d /= md;
    }
    if (c) out.y = d; else out.x = d;
  }
  st2m(FLD(dss), ix, out, m0, m1);
}
