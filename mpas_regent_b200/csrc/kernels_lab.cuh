// kernels_lab.cuh -- laboratory variants kept for profiles/ scripts only.  Compiled into the library ONLY with
// -DMPASB200_LAB (make lab); the shipped libmpas_b200.so does not contain them.
#pragma once
// ============================================================================================
// EXPERIMENTAL variants of k_divdamp used to measure which latency-hiding structure pays on B200
// (profiles/r1_divdamp_variants.md).  Selected through mpasb200_debug_divdamp only.
// V1: skip flag and ecv fetched together, own-column load issued before the dependent gathers
__global__ void k_divdamp_v1(const View V, double coef_divdamp) {
  PAIR_THREAD(V.nEdges)
  if (!m0) return;
  const unsigned char skip = V.divdampSkip[x];
  const int4 cv = V.ecv[x];
  const D2 r = ld2(FLD(ru_p), ix);
  const double sz = 1.0 - V.specZoneMaskEdge[x];
  if (skip) return;
  const double* rpp = FLD(rtheta_pp); const double* rppo = FLD(rtheta_pp_old); const double* tm = FLD(theta_m);
  const D2 a1 = G2(rpp, cv.x), b1 = G2(rppo, cv.x), a2 = G2(rpp, cv.y), b2 = G2(rppo, cv.y), t1 = G2(tm, cv.x), t2 = G2(tm, cv.y);
  st2m(FLD(ru_p), ix, r + coef_divdamp * ((-(a2 - b2)) - (-(a1 - b1))) * sz / (t1 + t2), m0, m1);
}
// V2: V1 + every block prefetches the index words of the block that will run ~one wave later into L2
__global__ void k_divdamp_v2(const View V, double coef_divdamp, int ahead) {
  PAIR_THREAD(V.nEdges)
  if (threadIdx.x == 0) {
    const long xa = (long)x + (long)ahead * blockDim.y;
    if (xa < V.nEdges) { prefetch_l2(&V.ecv[xa]); prefetch_l2(&V.divdampSkip[xa]); prefetch_l2(&V.specZoneMaskEdge[xa]); }
  }
  if (!m0) return;
  const unsigned char skip = V.divdampSkip[x];
  const int4 cv = V.ecv[x];
  const D2 r = ld2(FLD(ru_p), ix);
  const double sz = 1.0 - V.specZoneMaskEdge[x];
  if (skip) return;
  const double* rpp = FLD(rtheta_pp); const double* rppo = FLD(rtheta_pp_old); const double* tm = FLD(theta_m);
  const D2 a1 = G2(rpp, cv.x), b1 = G2(rppo, cv.x), a2 = G2(rpp, cv.y), b2 = G2(rppo, cv.y), t1 = G2(tm, cv.x), t2 = G2(tm, cv.y);
  st2m(FLD(ru_p), ix, r + coef_divdamp * ((-(a2 - b2)) - (-(a1 - b1))) * sz / (t1 + t2), m0, m1);
}
// V3: two edges per thread (x and x + half), all 14 gathers in flight together
__global__ void k_divdamp_v3(const View V, double coef_divdamp) {
  const int k0 = 2 * (int)threadIdx.x, k1 = k0 + 1;
  const int LP = V.LP, L = V.L;
  const int half = (V.nEdges + 1) / 2;
  const int xa = blockIdx.x * blockDim.y + threadIdx.y, xb = xa + half;
  const bool ina = xa < half && k0 < L, inb = xb < V.nEdges && xa < half && k0 < L;
  if (!ina) return;
  const bool m1 = k1 < L;
  const size_t ia = (size_t)xa * LP + k0, ib = (size_t)(inb ? xb : xa) * LP + k0;
  const unsigned char sa = V.divdampSkip[xa], sb = inb ? V.divdampSkip[xb] : 1;
  const int4 ca = V.ecv[xa], cb = V.ecv[inb ? xb : xa];
  const D2 ra = ld2(FLD(ru_p), ia), rb = ld2(FLD(ru_p), ib);
  const double za = 1.0 - V.specZoneMaskEdge[xa], zb = 1.0 - V.specZoneMaskEdge[inb ? xb : xa];
  const double* rpp = FLD(rtheta_pp); const double* rppo = FLD(rtheta_pp_old); const double* tm = FLD(theta_m);
  const D2 a1 = G2(rpp, ca.x), b1 = G2(rppo, ca.x), a2 = G2(rpp, ca.y), b2 = G2(rppo, ca.y), t1 = G2(tm, ca.x), t2 = G2(tm, ca.y);
  const D2 c1 = G2(rpp, cb.x), d1 = G2(rppo, cb.x), c2 = G2(rpp, cb.y), d2 = G2(rppo, cb.y), u1 = G2(tm, cb.x), u2 = G2(tm, cb.y);
  if (!sa) st2m(FLD(ru_p), ia, ra + coef_divdamp * ((-(a2 - b2)) - (-(a1 - b1))) * za / (t1 + t2), true, m1);
  if (!sb) st2m(FLD(ru_p), ib, rb + coef_divdamp * ((-(c2 - d2)) - (-(c1 - d1))) * zb / (u1 + u2), true, m1);
}
// V4: persistent blocks looping over edge tiles, the next tile's index words are loaded before the
// current tile's gathers are consumed
__global__ void k_divdamp_v4(const View V, double coef_divdamp) {
  const int k0 = 2 * (int)threadIdx.x, k1 = k0 + 1;
  const int LP = V.LP, L = V.L;
  if (k0 >= L) return;
  const bool m1 = k1 < L;
  const int stride = gridDim.x * blockDim.y;
  int x = blockIdx.x * blockDim.y + threadIdx.y;
  if (x >= V.nEdges) return;
  const double* rpp = FLD(rtheta_pp); const double* rppo = FLD(rtheta_pp_old); const double* tm = FLD(theta_m);
  int4 cv = V.ecv[x]; unsigned char skip = V.divdampSkip[x]; double sz = 1.0 - V.specZoneMaskEdge[x];
  D2 r = ld2(FLD(ru_p), (size_t)x * LP + k0);
  while (true) {
    const int xn = x + stride;
    const bool more = xn < V.nEdges;
    const int xs = more ? xn : x;
    const int4 cvn = V.ecv[xs]; const unsigned char skn = V.divdampSkip[xs]; const double szn = 1.0 - V.specZoneMaskEdge[xs];
    const D2 rn = ld2(FLD(ru_p), (size_t)xs * LP + k0);
    if (!skip) {
      const D2 a1 = G2(rpp, cv.x), b1 = G2(rppo, cv.x), a2 = G2(rpp, cv.y), b2 = G2(rppo, cv.y), t1 = G2(tm, cv.x), t2 = G2(tm, cv.y);
      st2m(FLD(ru_p), (size_t)x * LP + k0, r + coef_divdamp * ((-(a2 - b2)) - (-(a1 - b1))) * sz / (t1 + t2), true, m1);
    }
    if (!more) break;
    x = xn; cv = cvn; skip = skn; sz = szn; r = rn;
  }
}
// V5: four levels per thread (two 128-bit words), half the threads per column
__global__ void k_divdamp_v5(const View V, double coef_divdamp) {
  const int k0 = 4 * (int)threadIdx.x;
  const int LP = V.LP, L = V.L;
  const int x = blockIdx.x * blockDim.y + threadIdx.y;
  if (x >= V.nEdges || k0 >= L) return;
  const size_t ix = (size_t)x * LP + k0;
  const unsigned char skip = V.divdampSkip[x];
  const int4 cv = V.ecv[x];
  const D2 r0 = ld2(FLD(ru_p), ix), r1 = ld2(FLD(ru_p), ix + 2);
  const double sz = 1.0 - V.specZoneMaskEdge[x];
  if (skip) return;
  const double* rpp = FLD(rtheta_pp); const double* rppo = FLD(rtheta_pp_old); const double* tm = FLD(theta_m);
  const size_t i1 = (size_t)cv.x * LP + k0, i2 = (size_t)cv.y * LP + k0;
  const D2 a1 = ld2(rpp, i1), b1 = ld2(rppo, i1), a2 = ld2(rpp, i2), b2 = ld2(rppo, i2), t1 = ld2(tm, i1), t2 = ld2(tm, i2);
  const D2 A1 = ld2(rpp, i1 + 2), B1 = ld2(rppo, i1 + 2), A2 = ld2(rpp, i2 + 2), B2 = ld2(rppo, i2 + 2), T1 = ld2(tm, i1 + 2), T2 = ld2(tm, i2 + 2);
  st2m(FLD(ru_p), ix, r0 + coef_divdamp * ((-(a2 - b2)) - (-(a1 - b1))) * sz / (t1 + t2), true, k0 + 1 < L);
  st2m(FLD(ru_p), ix + 2, r1 + coef_divdamp * ((-(A2 - B2)) - (-(A1 - B1))) * sz / (T1 + T2), k0 + 2 < L, k0 + 3 < L);
}
// V6: one-wave-ahead L2 prefetch of DATA as well as index words.  Exact: the index word of the tile `ahead`
// blocks later (itself prefetched 2*ahead earlier) is loaded and the gather lines it points to are prefetched.
DI void prefetch_col(const double* p, size_t col, int LP, int k0) { prefetch_l2(p + col * LP + k0); }
__global__ void k_divdamp_v6(const View V, double coef_divdamp, int ahead) {
  PAIR_THREAD(V.nEdges)
  const bool pf_lane = (threadIdx.x & 7) == 0;          // one lane per 128-byte line of a column
  const long xa = (long)x + (long)ahead * blockDim.y, xb = (long)x + 2L * ahead * blockDim.y;
  unsigned char skip = 1; int4 cv = make_int4(0, 0, 0, 0); D2 r = bc(0); double sz = 0;
  if (m0) {
    skip = V.divdampSkip[x]; cv = V.ecv[x]; r = ld2(FLD(ru_p), ix); sz = 1.0 - V.specZoneMaskEdge[x];
  }
  if (pf_lane && k0 < V.L) {
    if (xb < V.nEdges && threadIdx.x == 0) { prefetch_l2(&V.ecv[xb]); prefetch_l2(&V.divdampSkip[xb]); prefetch_l2(&V.specZoneMaskEdge[xb]); }
    if (xa < V.nEdges) {
      const int4 ca = V.ecv[xa];
      prefetch_col(FLD(ru_p), xa, LP, k0);
      prefetch_col(FLD(rtheta_pp), ca.x, LP, k0); prefetch_col(FLD(rtheta_pp_old), ca.x, LP, k0); prefetch_col(FLD(theta_m), ca.x, LP, k0);
      prefetch_col(FLD(rtheta_pp), ca.y, LP, k0); prefetch_col(FLD(rtheta_pp_old), ca.y, LP, k0); prefetch_col(FLD(theta_m), ca.y, LP, k0);
    }
  }
  if (!m0 || skip) return;
  const double* rpp = FLD(rtheta_pp); const double* rppo = FLD(rtheta_pp_old); const double* tm = FLD(theta_m);
  const D2 a1 = G2(rpp, cv.x), b1 = G2(rppo, cv.x), a2 = G2(rpp, cv.y), b2 = G2(rppo, cv.y), t1 = G2(tm, cv.x), t2 = G2(tm, cv.y);
  st2m(FLD(ru_p), ix, r + coef_divdamp * ((-(a2 - b2)) - (-(a1 - b1))) * sz / (t1 + t2), m0, m1);
}
// V7: V4 (persistent, next tile's index words in registers) with four levels per thread
__global__ void k_divdamp_v7(const View V, double coef_divdamp) {
  const int k0 = 4 * (int)threadIdx.x;
  const int LP = V.LP, L = V.L;
  if (k0 >= L) return;
  const int stride = gridDim.x * blockDim.y;
  int x = blockIdx.x * blockDim.y + threadIdx.y;
  if (x >= V.nEdges) return;
  const double* rpp = FLD(rtheta_pp); const double* rppo = FLD(rtheta_pp_old); const double* tm = FLD(theta_m);
  int4 cv = V.ecv[x]; unsigned char skip = V.divdampSkip[x]; double sz = 1.0 - V.specZoneMaskEdge[x];
  D2 r0 = ld2(FLD(ru_p), (size_t)x * LP + k0), r1 = ld2(FLD(ru_p), (size_t)x * LP + k0 + 2);
  while (true) {
    const int xn = x + stride;
    const bool more = xn < V.nEdges;
    const int xs = more ? xn : x;
    const int4 cvn = V.ecv[xs]; const unsigned char skn = V.divdampSkip[xs]; const double szn = 1.0 - V.specZoneMaskEdge[xs];
    const D2 rn0 = ld2(FLD(ru_p), (size_t)xs * LP + k0), rn1 = ld2(FLD(ru_p), (size_t)xs * LP + k0 + 2);
    if (!skip) {
      const size_t i1 = (size_t)cv.x * LP + k0, i2 = (size_t)cv.y * LP + k0, ix = (size_t)x * LP + k0;
      const D2 a1 = ld2(rpp, i1), b1 = ld2(rppo, i1), a2 = ld2(rpp, i2), b2 = ld2(rppo, i2), t1 = ld2(tm, i1), t2 = ld2(tm, i2);
      const D2 A1 = ld2(rpp, i1 + 2), B1 = ld2(rppo, i1 + 2), A2 = ld2(rpp, i2 + 2), B2 = ld2(rppo, i2 + 2), T1 = ld2(tm, i1 + 2), T2 = ld2(tm, i2 + 2);
      st2m(FLD(ru_p), ix, r0 + coef_divdamp * ((-(a2 - b2)) - (-(a1 - b1))) * sz / (t1 + t2), true, k0 + 1 < L);
      st2m(FLD(ru_p), ix + 2, r1 + coef_divdamp * ((-(A2 - B2)) - (-(A1 - B1))) * sz / (T1 + T2), k0 + 2 < L, k0 + 3 < L);
    }
    if (!more) break;
    x = xn; cv = cvn; skip = skn; sz = szn; r0 = rn0; r1 = rn1;
  }
}
// V8: persistent tile loop WITHOUT next-tile prefetch (isolates the effect of block scheduling overhead)
__global__ void k_divdamp_v8(const View V, double coef_divdamp) {
  const int k0 = 2 * (int)threadIdx.x, k1 = k0 + 1;
  const int LP = V.LP, L = V.L;
  if (k0 >= L) return;
  const bool m1 = k1 < L;
  const int stride = gridDim.x * blockDim.y;
  const double* rpp = FLD(rtheta_pp); const double* rppo = FLD(rtheta_pp_old); const double* tm = FLD(theta_m);
  for (int x = blockIdx.x * blockDim.y + threadIdx.y; x < V.nEdges; x += stride) {
    if (V.divdampSkip[x]) continue;
    const int4 cv = V.ecv[x];
    const size_t ix = (size_t)x * LP + k0;
    const D2 r = ld2(FLD(ru_p), ix);
    const double sz = 1.0 - V.specZoneMaskEdge[x];
    const D2 a1 = G2(rpp, cv.x), b1 = G2(rppo, cv.x), a2 = G2(rpp, cv.y), b2 = G2(rppo, cv.y), t1 = G2(tm, cv.x), t2 = G2(tm, cv.y);
    st2m(FLD(ru_p), ix, r + coef_divdamp * ((-(a2 - b2)) - (-(a1 - b1))) * sz / (t1 + t2), true, m1);
  }
}
