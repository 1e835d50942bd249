// view.h -- device-side view of the structure-of-arrays mirror (passed to kernels by value).
//
// Layout in HBM (DESIGN.md "Data layout"):
//   3-D field            : double[(N+1)][LP]        levels contiguous, LP = roundup(nVertLevels+1, 4)
//                          (every column starts on a 32-byte sector, level pairs are 16-byte aligned
//                          -> 128-bit loads), entity N is the zero pad entity
//   array-typed 3-D field: double[slots][(N+1)][LP] (slot-major: each slot is its own 3-D field)
//   static per-entity    : T[(N+1)][width]          resolved 0-based ids, pad row N points at pads
//   per-(cell, slot)     : copies of the per-edge statics a cell needs for its edgesOnCell slots
//                          (dvEdge, invDcEdge, cellsOnEdge, meshScaling, advection lists), so a
//                          cell kernel reaches its neighbour columns after ONE level of index loads
// Entities are stored in space-filling-curve order; the permutation lives on the host side of the
// library and in perm arrays used only by upload/download/pack/unpack.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector_types.h>
#include "../../include/mpas_b200.h"

struct View {
  // dimensions
  int nCells, nEdges, nVertices, L, LP;
  int xoff, xend;    // a launch covers internal indices [xoff, min(n, xend)) of its entity type (mpasb200_set_range)
  int maxEdges, maxEdges2, vertexDegree, nAdv;
  int MEP;           // edgesOnCell row pitch (ints), multiple of 4
  int NAP;           // advection list pitch (entries per (cell, slot)), even
  size_t cellSlot;   // (nCells+1)*LP: distance between slots of a cell array-typed field
  // 3-D fields + vertical fields, by field id
  double* f[MPASB200_F_COUNT];
  // cell statics
  const int* nEdgesOnCell; const int* edgesOnCell /* [c][MEP] */; const int* verticesOnCell; const int* kiteForCell;
  const int* c1OnCell; const int* c2OnCell;            // [c][MEP]: cellsOnEdge[e][0/1] of the cell's slot-i edge
  const double* edgesOnCellSign; const double* edgesOnCell_sign; const double* invAreaCell; const double* cosLatCell;
  const double* dvOnCell; const double* invDcOnCell; const double* ms2OnCell; const double* ms4OnCell;   // [c][maxEdges]
  const double* dcOnCell;      // [c][maxEdges]  dcEdge of the slot-i edge
  const double* defc_a; const double* defc_b; const int* bdyMaskCell; const double* specZoneMaskCell;
  const unsigned char* isShared; const unsigned char* inCpr;
  const double* sinLatCell; const double* cosLonCell; const double* sinLonCell;   // host-evaluated (glibc), like cosLatCell
  const double* coeffsRecon;   // [c][maxEdges][3]  coeffs_reconstruct
  const int* nAdvOnCell;       // [c][maxEdges]
  const int* advCellOnCell;    // [c][maxEdges][NAP]
  const double* advCoefOnCell; const double* adv3OnCell;   // [c][maxEdges][NAP]
  int advFinite;               // every advection coefficient is finite and small enough that coef + coef_3rd cannot overflow (upload_mesh)
  // edge statics
  const int4* ecv;             // {cell1, cell2, vertex1, vertex2}
  const int* cellsOnEdge; const int* verticesOnEdge; const int* nEdgesOnEdge; const int* edgesOnEdge_ECP; const int* edgesOnEdge;
  const double* weightsOnEdge; const double* dcEdge; const double* dvEdge; const double* invDcEdge; const double* invDvEdge;
  const double* cosAngleEdge; const double* sinAngleEdge; const double* cosLatEdge;
  const int* nAdvCellsForEdge; const int* advCellsForEdge; const double* adv_coefs; const double* adv_coefs_3rd;
  int NAE; const int* advCellE /* [e][NAE] */; const double2* advCoefE /* [e][NAE] {adv_coefs, adv_coefs_3rd} */;   // 16-byte rows for k_dt_theta_flux
  const double* meshScalingDel2; const double* meshScalingDel4; const double* specZoneMaskEdge;
  const unsigned char* divdampSkip;   // isShared[cell1] && isShared[cell2]   (dynamics_tasks.rg:1750)
  // vertex statics
  const int* edgesOnVertex; const double* edgesOnVertexSign; const double* edgesOnVertex_sign; const double* kiteAreasOnVertex;
  const double* dcOnVertex;    // [v][3]: dcEdge of the vertex's edges
  const double* fVertex; const double* invAreaTriangle;
  // scratch (library-private, not region fields)
  double* scr_flux;                   // horizontal theta flux per edge (k_dt_theta_flux)  [(nEdges+1)][LP]
  double* scr_rs; double* scr_ts;     // horizontal flux parts of rs/ts in the two-kernel acoustic step  [(nCells+1)][LP]
};

// tile lists of k_dt_edge_tile (kernels_tiles.cuh, laboratory builds); the handle holds one in every build so that its layout does not depend on the build
struct EdgeTiles {
  const int* cols;             // [tile][cap]  internal edge ids of the staged columns; the tile's own edges come first (slot = local index)
  const int* ncols;            // [tile]
  const unsigned char* slot;   // [edge][SP]   byte j = staged slot of edgesOnEdge[j], 255 = not staged (read from global memory)
  int cap, SP, TE;
};

// scalar constants a task needs (filled on the host from MpasConfig + task arguments)
struct DynTendParams {
  int rk_step; int mixing; int mix_full; int rayleigh_u; int visc4_on; int cam_on; int vmix_u_on; int vmix_t_on;
  double kdiff_scale;      // (c_s*len_disp)^2
  double kdiff_cap;        // 0.01*len_disp^2 * (1/dt)
  double cam_base;         // unused (kept for layout stability)
  double h_mom_eddy_visc4, h_theta_eddy_visc4, v_mom_eddy_visc2, v_theta_eddy_visc2;
  double prandtl_inv, r_earth, inv_r_earth, omega2 /* 2*omega */, gravity, del4u_div_factor;
  double rayleigh_coef_inverse; int rayleigh_levels;
};
