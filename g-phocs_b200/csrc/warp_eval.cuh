// warp_eval.cuh — computeLocusDataLikelihood(locus, useOld = 1) (LocusDataLikelihood.c:426-483, 1559-1673) for ONE
// locus by ONE warp, for kernels that keep a locus on a warp across many proposals (k_smp_sweep).  Same arithmetic,
// same summation order and the same buffer-flip protocol as k_eval (clv_kernels.cuh), so the two paths leave
// bit-identical conditional vectors and log-likelihoods; genealogies of up to 64 nodes.
//
// Lanes are nodes while the work list is built and columns while vectors are computed: a column is always handled
// by the same lane, so a recomputed child is read back by the lane that wrote it.
#pragma once
#include "clv_kernels.cuh"

namespace gphocs {

__device__ inline void warpChildVector(const StoreDev& d, const double* __restrict__ clvLocus, int child, uint32_t sel, int P, int p,
                                       long long col, double (&v)[4]) {
  if (child < d.n) {   // leaf: 4-bit base mask -> 0/1 conditionals (computeLeafConditionals, .c:1336-1386)
    const unsigned mask = (unsigned)(d.leafWords[(size_t)(child >> 4) * d.Ct + col] >> ((child & 15) * 4)) & 15u;
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] = (mask >> q) & 1u ? 1.0 : 0.0;
  } else {
    const double2* g = reinterpret_cast<const double2*>(clvLocus + ((size_t)((child - d.n) * 2 + sel) * P + p) * 4);
    const double2 x = g[0], y = g[1];
    v[0] = x.x; v[1] = x.y; v[2] = y.x; v[3] = y.y;
  }
}

__device__ inline void warpEvalIncremental(const StoreDev& d, const TreeView& t, int l, int lane) {
  const int n = d.n, N = d.N;
  const int c0 = d.colStart[l];
  const int P = d.colStart[l + 1] - c0;
  const int root = *t.root;
  if (P <= 0 || root < n) return;
  if (lane == 0) *t.savedLnL = *t.lnL;   // always, even when nothing is recomputed (.c:440)
  // ---- dirty nodes and their ancestors (the recursion condition of computeConditionalJC_new, .c:1583); a moved
  //      leaf dirties its father (.c:1569-1575)
  unsigned long long dirty = 0ull, need = 0ull;
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int x = lane + 32 * r;
    const bool f = x < N && (t.node[x].flags & F_RECALC);
    dirty |= (unsigned long long)__ballot_sync(0xffffffffu, f) << (32 * r);
  }
  for (unsigned long long m = dirty; m; m &= m - 1) {
    const int v = __ffsll((long long)m) - 1;
    int u = v < n ? t.node[v].father : v;
    while (u >= 0 && !((need >> u) & 1ull)) {
      need |= 1ull << u;
      u = t.node[u].father;
    }
  }
  if (!need) return;
  // ---- marked nodes write into their other buffer (copyNodeConditionals, .c:1889-1906: once per proposal)
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int x = lane + 32 * r;
    if (x < N && ((need >> x) & 1ull)) {
      const uint8_t f = t.node[x].flags;
      if (!(f & F_RECALC)) t.node[x].flags = (uint8_t)((f ^ F_SEL) | F_RECALC);
    }
  }
  __syncwarp();
  const double rate = *t.rate;
  double* clvLocus = d.clv + (size_t)c0 * d.NI * 8;
  unsigned long long done = 0ull;
  while (done != need) {
    // nodes whose marked children have been recomputed
    unsigned long long ready = 0ull;
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int x = lane + 32 * r;
      bool ok = false;
      if (x < N && ((need >> x) & 1ull) && !((done >> x) & 1ull)) {
        const NodeRec rec = t.node[x];
        const bool lOk = !((need >> rec.left) & 1ull) || ((done >> rec.left) & 1ull);
        const bool rOk = !((need >> rec.right) & 1ull) || ((done >> rec.right) & 1ull);
        ok = lOk && rOk;
      }
      ready |= (unsigned long long)__ballot_sync(0xffffffffu, ok) << (32 * r);
    }
    for (unsigned long long m = ready; m; m &= m - 1) {
      const int v = __ffsll((long long)m) - 1;
      const NodeRec rv = t.node[v];
      const int A = rv.left, B = rv.right;
      const double av = t.age[v];
      const double e0A = edgeProb(rate * (av - t.age[A])), e1A = 1.0 - 4.0 * e0A;
      const double e0B = edgeProb(rate * (av - t.age[B])), e1B = 1.0 - 4.0 * e0B;
      const uint32_t selA = A >= n ? (t.node[A].flags & F_SEL) : 0u, selB = B >= n ? (t.node[B].flags & F_SEL) : 0u;
      double* dstRec = clvLocus + (size_t)((v - n) * 2 + (rv.flags & F_SEL)) * P * 4;
      for (int p = lane; p < P; p += 32) {
        double a[4], b[4], out[4];
        warpChildVector(d, clvLocus, A, selA, P, p, (long long)c0 + p, a);
        warpChildVector(d, clvLocus, B, selB, P, p, (long long)c0 + p, b);
        // computeSubtreeConditionals_new (.c:1650-1673) for both children
        const double sA = ((a[0] + a[1]) + a[2]) + a[3];
        const double sB = ((b[0] + b[1]) + b[2]) + b[3];
        const double qA = sA * e0A, qB = sB * e0B;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const double fa = sA >= 4.0 ? 1.0 : (qA + a[q] * e1A);   // an all-missing subtree contributes exactly 1
          const double fb = sB >= 4.0 ? 1.0 : (qB + b[q] * e1B);
          out[q] = fa * fb;
        }
        double2* dst = reinterpret_cast<double2*>(dstRec + (size_t)p * 4);
        dst[0] = make_double2(out[0], out[1]);
        dst[1] = make_double2(out[2], out[3]);
      }
    }
    done |= ready;
  }
  __syncwarp();
  // ---- root: sum over the 4*phases conditionals of each phase group, count * log, patterns added in order (.c:470-479)
  const double* rootRec = clvLocus + (size_t)((root - n) * 2 + (t.node[root].flags & F_SEL)) * P * 4;
  double lnl = 0.0;
  for (int p0 = 0; p0 < P; p0 += 32) {
    const int p = p0 + lane;
    double term = 0.0;
    if (p < P) {
      const int ph = d.grpPhases[c0 + p];
      if (ph > 0) {
        double prob = 0.0;
        const int numConds = 4 * ph;
        for (int j = 0; j < numConds; j++) prob += rootRec[(size_t)p * 4 + j];
        term = log(prob / numConds) * d.grpCount[c0 + p];
      }
    }
    const int m = min(32, P - p0);
    for (int j = 0; j < m; j++) lnl += __shfl_sync(0xffffffffu, term, j);
  }
  if (lane == 0) *t.lnL = lnl;
  __syncwarp();
}

}  // namespace gphocs
