// kernels_jw.cuh -- init_atm_case_jw on the device (reference vertical_init/init_atm_cases.rg:24-743).
//
// The reference routine cannot be restated literally (it indexes regions with swapped, far out-of-range (level, cell) pairs,
// init_atm_cases.rg:266,268,419,447; zw[k] = (k-1)*dz with 0-based k, :198; sh[0] = -1, :178).  These kernels are the CORRECTED
// reading -- what the MPAS Fortran does -- formula by formula in the operation order of mpas_regent_b200/init_jw.py, the host
// generator every parity test is fed by; tests compare the two at 1e-12 (exp / pow / sin / cos differ in the last place between
// libm and the device).  qv = 0 (dry case).  Constants: init_atm_cases.rg:46-75, constants.rg:27-38.
#pragma once
#include "kernels.cuh"

struct JwParams {
  int nlat;                      // rows of the latitude table
  double u0, t0, t0b, dtdz, eta_t, delta_t, etavs0, p0, zt, r_earth, omega, rgas, cp, gravity, pii;
  const double* sh; const double* ah; const double* dzw; const double* dzu;      // [L+1] vertical-grid helpers (device)
  double* pp_t; double* tt_t;    // [nlat][L] hydrostatically balanced columns of the (z, lat) section
  const double* latCell; const double* areaCell;   // [nCells+1], internal numbering
  const double* latVertex;                         // [nVertices+1]
};

DI double jw_terrain(const JwParams& J, double phi) {                       // :144-160
  const double ce = pow(cos(J.etavs0), 1.5);
  const double s = sin(phi), c = cos(phi);
  return J.u0 / J.gravity * ce * ((-2.0 * pow(s, 6.0) * (c * c + 1.0 / 3.0) + 10.0 / 63.0) * J.u0 * ce
                                  + (1.6 * pow(c, 3.0) * (s * s + 2.0 / 3.0) - J.pii / 4.0) * J.r_earth * J.omega);
}
DI double jw_zgrid(const JwParams& J, double hx, int k) {                    // :165-255 (terrain-following heights)
  return (1.0 - J.ah[k]) * (J.sh[k] * (J.zt - hx) + hx) + (J.ah[k] * J.sh[k]) * J.zt;
}

// the (z, lat) section: 10 x 25 hydrostatic iterations per column (:278-383, 417-516), one thread per table latitude.
// Scratch rows pp, tt, rr, zz, ppb, rb live in global memory ([6][nlat][L], `wk`).
__global__ void k_jw_table(const JwParams J, const View V, double* __restrict__ wk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= J.nlat) return;
  const int L = V.L;
  const size_t row = (size_t)i * L, plane = (size_t)J.nlat * L;
  double* pp = J.pp_t + row; double* tt = J.tt_t + row;
  double* rr = wk + row; double* zz = wk + plane + row; double* ppb = wk + 2 * plane + row; double* rb = wk + 3 * plane + row;
  const double step = (0.5 * J.pii - (-0.5 * J.pii)) / (double)(J.nlat - 1);
  const double phi = (i == J.nlat - 1) ? 0.5 * J.pii : -0.5 * J.pii + (double)i * step;      // numpy.linspace
  const double hx = jw_terrain(J, phi);
  const double* fzm = FLD(fzm); const double* fzp = FLD(fzp);
  for (int k = 0; k < L; ++k) {
    const double z0 = jw_zgrid(J, hx, k), z1 = jw_zgrid(J, hx, k + 1);
    zz[k] = J.dzw[k] / (z1 - z0);
    const double ztm = 0.5 * (z1 + z0);
    ppb[k] = J.p0 * exp(-J.gravity * ztm / (J.rgas * J.t0b));
    rb[k] = ppb[k] / (J.rgas * J.t0b * zz[k]);
    pp[k] = 0.0;
  }
  const double s = sin(phi), c = cos(phi);
  const double A = -2.0 * pow(s, 6.0) * (c * c + 1.0 / 3.0) + 10.0 / 63.0;
  const double B = (1.6 * pow(c, 3.0) * (s * s + 2.0 / 3.0) - J.pii / 4.0) * J.r_earth * J.omega;
  for (int itr = 0; itr < 10; ++itr) {
    for (int k = 0; k < L; ++k) {
      const double eta = (ppb[k] + pp[k]) / J.p0;
      const double etav = (eta - 0.252) * J.pii / 2.0;
      const double teta = J.t0 * pow(eta, J.rgas * J.dtdz / J.gravity) + ((eta >= J.eta_t) ? 0.0 : J.delta_t * pow(dmax(J.eta_t - eta, 0.0), 5.0));
      const double cosv = dmax(cos(etav), 0.0);
      tt[k] = teta + 0.75 * eta * J.pii * J.u0 / J.rgas * sin(etav) * sqrt(cosv) * (A * 2.0 * J.u0 * pow(cosv, 1.5) + B);
    }
    for (int itrp = 0; itrp < 25; ++itrp) {
      for (int k = 0; k < L; ++k) rr[k] = (pp[k] / (J.rgas * zz[k]) - rb[k] * (tt[k] - J.t0b)) / tt[k];
      const double ppi0 = J.p0 - 0.5 * J.dzw[0] * J.gravity * (1.25 * (rr[0] + rb[0]) - 0.25 * (rr[1] + rb[1])) - ppb[0];
      double cum = 0.0;
      pp[0] = 0.2 * ppi0 + 0.8 * pp[0];
      for (int k = 1; k < L; ++k) {
        const double term = (J.dzu[k] * J.gravity) * (rr[k - 1] * fzp[k] + rr[k] * fzm[k]);
        cum = (k == 1) ? term : cum + term;                                        // numpy.cumsum
        pp[k] = 0.2 * (ppi0 - cum) + 0.8 * pp[k];
      }
    }
  }
}

// cell columns: zgrid, zz, base state, the balanced pp / tt interpolated from the section, rho / theta (:417-522)
__global__ void k_jw_cell(const JwParams J, const View V) {
  PAIR_THREAD(V.nCells)
  if (!inx || k0 > L) return;
  const double lat = J.latCell[x];
  const double hx = jw_terrain(J, lat);
  const double fpos = (lat + 0.5 * J.pii) / (J.pii / (double)(J.nlat - 1));
  long i0 = (long)floor(fpos);
  i0 = i0 < 0 ? 0 : (i0 > J.nlat - 2 ? J.nlat - 2 : i0);
  const double wt = fpos - (double)i0;
  for (int c = 0; c < 2; ++c) {
    const int k = k0 + c;
    if (k > L) break;
    const size_t i = ix + c;
    const double zg = jw_zgrid(J, hx, k);
    FLD(zgrid)[i] = zg;
    double zz = 0.0, rb = 0.0, pp = 0.0, rr = 0.0, tb = 0.0, ex = 0.0, tm = 0.0, rtp = 0.0, rz = 0.0;
    if (k < L) {
      const double zg1 = jw_zgrid(J, hx, k + 1);
      zz = J.dzw[k] / (zg1 - zg);
      pp = (1.0 - wt) * J.pp_t[(size_t)i0 * L + k] + wt * J.pp_t[(size_t)(i0 + 1) * L + k];
      const double tt = (1.0 - wt) * J.tt_t[(size_t)i0 * L + k] + wt * J.tt_t[(size_t)(i0 + 1) * L + k];
      const double ztemp = 0.5 * (zg1 + zg);
      const double ppb = J.p0 * exp(-J.gravity * ztemp / (J.rgas * J.t0b));
      const double pb = pow(ppb / J.p0, J.rgas / J.cp);
      rb = ppb / (J.rgas * J.t0b * zz);
      tb = J.t0b / pb;
      rr = (pp / (J.rgas * zz) - rb * (tt - J.t0b)) / tt;
      ex = pow((ppb + pp) / J.p0, J.rgas / J.cp);
      tm = tt / ex;
      rtp = tm * rr + rb * (tm - tb);
      rz = rb + rr;
    }
    FLD(zz)[i] = zz; FLD(rho_base)[i] = rb; FLD(pressure_p)[i] = pp; FLD(rho_p)[i] = rr; FLD(theta_base)[i] = tb;
    FLD(exner)[i] = ex; FLD(theta_m)[i] = tm; FLD(rtheta_p)[i] = rtp; FLD(rho_zz)[i] = rz;
  }
}

// edges: zxu (:257-263), the zonal wind u and ru (:530-596), the metric terms zb (:616-665; zb3 = 0)
__global__ void k_jw_edge(const JwParams J, const View V) {
  PAIR_THREAD(V.nEdges)
  if (!inx || k0 > L) return;
  const int4 cv = V.ecv[x];
  const double dv = V.dvEdge[x], dc = V.dcEdge[x];
  const double lat1 = J.latVertex[cv.z], lat2 = J.latVertex[cv.w];
  const double flux = (0.5 * (lat2 - lat1) - 0.125 * (sin(4.0 * lat2) - sin(4.0 * lat1))) * J.r_earth / dv;
  const double* zg = FLD(zgrid); const double* ppf = FLD(pressure_p); const double* rz = FLD(rho_zz);
  const size_t edgeSlot = (size_t)(V.nEdges + 1) * LP;
  const double a1 = dv / J.areaCell[cv.x], a2 = dv / J.areaCell[cv.y];
  for (int c = 0; c < 2; ++c) {
    const int k = k0 + c;
    if (k > L) break;
    const size_t i = ix + c, i1 = (size_t)cv.x * LP + k, i2 = (size_t)cv.y * LP + k;
    double zxu = 0.0, u = 0.0, ru = 0.0, zb0 = 0.0, zb1 = 0.0;
    if (k < L) {
      zxu = 0.5 * (zg[i2] - zg[i1] + zg[i2 + 1] - zg[i1 + 1]) / dc;
      const double pb1 = J.p0 * exp(-J.gravity * (0.5 * (zg[i1 + 1] + zg[i1])) / (J.rgas * J.t0b));
      const double pb2 = J.p0 * exp(-J.gravity * (0.5 * (zg[i2 + 1] + zg[i2])) / (J.rgas * J.t0b));
      const double etavs = (0.5 * ((pb1 + ppf[i1]) + (pb2 + ppf[i2])) / J.p0 - 0.252) * J.pii / 2.0;
      u = J.u0 * flux * pow(dmax(cos(etavs), 0.0), 1.5);
      ru = 0.5 * (rz[i1] + rz[i2]) * u;
      const double z_edge = 0.5 * (zg[i1] + zg[i2]);
      zb0 = (z_edge - zg[i1]) * a1;
      zb1 = (z_edge - zg[i2]) * a2;
    }
    FLD(zxu)[i] = zxu; FLD(u)[i] = u; FLD(ru)[i] = ru;
    FLD(zb)[i] = zb0; FLD(zb)[edgeSlot + i] = zb1;
    FLD(zb3)[i] = 0.0; FLD(zb3)[edgeSlot + i] = 0.0;
  }
}

// rw and w (:681-704): the edge contributions gathered per cell in edgesOnCell slot order (deterministic; the host generator
// scatters them edge by edge -- same terms, another summation order)
__global__ void k_jw_rw(const View V) {
  PAIR_THREAD(V.nCells)
  if (!inx || k0 > L) return;
  const double* fzm = FLD(fzm); const double* fzp = FLD(fzp);
  const double* zz = FLD(zz); const double* ru = FLD(ru); const double* rz = FLD(rho_zz); const double* zb = FLD(zb);
  const size_t edgeSlot = (size_t)(V.nEdges + 1) * LP;
  const int n = V.nEdgesOnCell[x];
  for (int c = 0; c < 2; ++c) {
    const int k = k0 + c;
    if (k > L) break;
    const size_t i = ix + c;
    double rw = 0.0, w = 0.0;
    if (k >= 1 && k < L) {
      const double zzf = fzm[k] * zz[i] + fzp[k] * zz[i - 1];
      for (int s = 0; s < n; ++s) {
        const int e = V.edgesOnCell[x * V.MEP + s];
        const size_t ie = (size_t)e * LP + k;
        const double fl = fzm[k] * ru[ie] + fzp[k] * ru[ie - 1];
        if (x == V.c1OnCell[x * V.MEP + s]) rw += -zzf * zb[ie] * fl;
        else rw += zzf * zb[edgeSlot + ie] * fl;
      }
      w = rw / (fzp[k] * rz[i - 1] + fzm[k] * rz[i]);
    }
    FLD(rw)[i] = rw; FLD(w)[i] = w;
  }
}
