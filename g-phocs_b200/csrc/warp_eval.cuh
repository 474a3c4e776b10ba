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

// ------------------------------------------------------------------------------------------------------------------
// Schedule hand-over.  A proposal kernel that already has the locus on a warp can finish the tree-side half of an
// incremental evaluation itself — mark dirty nodes and ancestors, flip their buffers, order them children first,
// compute the JC69 edge terms — and leave a compact schedule; k_eval_sched then only walks columns (phases E and F of
// k_eval), one thread per column over the store's CTA batches.  Same arithmetic as k_eval, bit for bit.
struct __align__(16) IncEntry {
  double e0A, e0B;       // edge terms p of the two children; 1 - 4p is recomputed by the consumer
  uint32_t a, b;         // child: leaf id | 0x80000000, or record index (node - n) * 2 + buffer
  uint32_t dst;          // record index of the node itself
  uint32_t pad;
};
static_assert(sizeof(IncEntry) == 32, "schedule entry layout");

// returns the number of entries written to out[0..NI) (uniform over the warp); the root's entry is the last one.
// Every lane loads the records and ages of its own nodes once (two coalesced loads); everything else — ancestor
// walks, children's buffer selectors and ages — comes from the other lanes' registers by shuffle.
__device__ inline int warpBuildSchedule(const StoreDev& d, const TreeView& t, int l, int lane, IncEntry* __restrict__ out) {
  const int n = d.n, N = d.N;
  const int P = d.colStart[l + 1] - d.colStart[l];
  const int root = *t.root;
  if (P <= 0 || root < n) return 0;
  if (lane == 0) *t.savedLnL = *t.lnL;   // always, even when nothing is recomputed (.c:440)
  NodeRec rec[2];
  double age[2];
  unsigned long long dirty = 0ull, need = 0ull;
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int x = lane + 32 * r;
    rec[r] = NodeRec{-1, -1, -1, 0, 0};
    age[r] = 0.0;
    if (x < N) { rec[r] = t.node[x]; age[r] = t.age[x]; }
    dirty |= (unsigned long long)__ballot_sync(0xffffffffu, (rec[r].flags & F_RECALC) != 0) << (32 * r);
  }
  auto fatherOf = [&](int u) {   // uniform u
    const int f0 = __shfl_sync(0xffffffffu, (int)rec[0].father, u & 31);
    const int f1 = __shfl_sync(0xffffffffu, (int)rec[1].father, u & 31);
    return u < 32 ? f0 : f1;
  };
  for (unsigned long long m = dirty; m; m &= m - 1) {
    const int v = __ffsll((long long)m) - 1;
    int u = v < n ? fatherOf(v) : v;   // a moved leaf dirties its father (.c:1569-1575)
    while (u >= 0 && !((need >> u) & 1ull)) {
      need |= 1ull << u;
      u = fatherOf(u);
    }
  }
  if (!need) return 0;
  // marked nodes write into their other buffer (copyNodeConditionals, .c:1889-1906: once per proposal)
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int x = lane + 32 * r;
    if (x < N && ((need >> x) & 1ull) && !(rec[r].flags & F_RECALC)) {
      rec[r].flags = (uint8_t)((rec[r].flags ^ F_SEL) | F_RECALC);
      t.node[x].flags = rec[r].flags;
    }
  }
  // position of every marked node in a children-first order: level by level from the bottom
  int pos[2] = {-1, -1};
  int base = 0;
  unsigned long long done = 0ull;
  while (done != need) {
    unsigned long long ready = 0ull;
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int x = lane + 32 * r;
      bool ok = false;
      if (x < N && ((need >> x) & 1ull) && !((done >> x) & 1ull)) {
        const bool lOk = !((need >> rec[r].left) & 1ull) || ((done >> rec[r].left) & 1ull);
        const bool rOk = !((need >> rec[r].right) & 1ull) || ((done >> rec[r].right) & 1ull);
        ok = lOk && rOk;
      }
      ready |= (unsigned long long)__ballot_sync(0xffffffffu, ok) << (32 * r);
    }
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int x = lane + 32 * r;
      if ((ready >> x) & 1ull) pos[r] = base + __popcll(ready & ((1ull << x) - 1ull));
    }
    base += __popcll(ready);
    done |= ready;
  }
  // every marked node writes its own entry; children's ages and buffer selectors by shuffle (lane-dependent sources)
  const double rate = *t.rate;
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int v = lane + 32 * r;
    const bool mine = pos[r] >= 0;
    const int A = mine ? rec[r].left : 0, B = mine ? rec[r].right : 0;
    const double ageA0 = __shfl_sync(0xffffffffu, age[0], A & 31), ageA1 = __shfl_sync(0xffffffffu, age[1], A & 31);
    const double ageB0 = __shfl_sync(0xffffffffu, age[0], B & 31), ageB1 = __shfl_sync(0xffffffffu, age[1], B & 31);
    const int flA0 = __shfl_sync(0xffffffffu, (int)rec[0].flags, A & 31), flA1 = __shfl_sync(0xffffffffu, (int)rec[1].flags, A & 31);
    const int flB0 = __shfl_sync(0xffffffffu, (int)rec[0].flags, B & 31), flB1 = __shfl_sync(0xffffffffu, (int)rec[1].flags, B & 31);
    if (!mine) continue;
    const double ageA = A < 32 ? ageA0 : ageA1, ageB = B < 32 ? ageB0 : ageB1;
    const int flA = A < 32 ? flA0 : flA1, flB = B < 32 ? flB0 : flB1;
    IncEntry en;
    en.e0A = edgeProb(rate * (age[r] - ageA));
    en.e0B = edgeProb(rate * (age[r] - ageB));
    en.a = A < n ? ((uint32_t)A | 0x80000000u) : (uint32_t)((A - n) * 2 + (flA & F_SEL));
    en.b = B < n ? ((uint32_t)B | 0x80000000u) : (uint32_t)((B - n) * 2 + (flB & F_SEL));
    en.dst = (uint32_t)((v - n) * 2 + (rec[r].flags & F_SEL));
    en.pad = 0;
    out[pos[r]] = en;
  }
  return base;
}

__device__ __forceinline__ void schedChildVector(const StoreDev& d, const double* __restrict__ clvLocus, uint32_t ref, int P, int p,
                                                 long long col, double (&v)[4]) {
  if (ref & 0x80000000u) {
    const int leaf = (int)(ref & 0x7fffffffu);
    const unsigned mask = (unsigned)(d.leafWords[(size_t)(leaf >> 4) * d.Ct + col] >> ((leaf & 15) * 4)) & 15u;
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] = (mask >> q) & 1u ? 1.0 : 0.0;
  } else {
    const double2* g = reinterpret_cast<const double2*>(clvLocus + ((size_t)ref * P + p) * 4);
    const double2 x = g[0], y = g[1];
    v[0] = x.x; v[1] = x.y; v[2] = y.x; v[3] = y.y;
  }
}

// column walk and root step of an incremental evaluation whose schedules were left by warpBuildSchedule: one thread
// per column over the store's CTA batches (batches of ordinary loci only: at most kThreads columns)
__global__ void __launch_bounds__(kThreads)
k_eval_sched(StoreDev d, const Batch* __restrict__ batches, const IncEntry* __restrict__ sched, const int* __restrict__ schedCount) {
  __shared__ double sRoot[kThreads * 4];
  __shared__ double sTerm[kThreads];
  __shared__ int mColStart[kMaxBatchLoci], mP[kMaxBatchLoci], mK[kMaxBatchLoci];
  const Batch b = batches[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nl = b.numLoci;
  int ph = 0, cnt = 0;
  const bool live = tid < b.numCols;
  const int c = b.firstCol + tid;
  if (live) { ph = d.grpPhases[c]; cnt = d.grpCount[c]; }
  if (tid < nl) {
    const int l = b.firstLocus + tid;
    const int c0 = d.colStart[l];
    mColStart[tid] = c0;
    mP[tid] = d.colStart[l + 1] - c0;
    mK[tid] = schedCount[l];
  }
  __syncthreads();
  int s = 0;
  if (live)
    while (s + 1 < nl && c >= mColStart[s + 1]) s++;
  const int k = live ? mK[s] : 0;
  double pv[4] = {0.0, 0.0, 0.0, 0.0};
  if (k > 0) {
    const int P = mP[s], p = c - mColStart[s];
    double* clvLocus = d.clv + (size_t)mColStart[s] * d.NI * 8;
    const IncEntry* en = sched + (size_t)(b.firstLocus + s) * d.NI;
    for (int e = 0; e < k; e++) {
      const double2 ee = *reinterpret_cast<const double2*>(&en[e].e0A);
      const uint4 ix = *reinterpret_cast<const uint4*>(&en[e].a);
      const double e1A = 1.0 - 4.0 * ee.x, e1B = 1.0 - 4.0 * ee.y;
      double a[4], bb[4];
      schedChildVector(d, clvLocus, ix.x, P, p, c, a);
      schedChildVector(d, clvLocus, ix.y, P, p, c, bb);
      // computeSubtreeConditionals_new (.c:1650-1673) for both children
      const double sA = ((a[0] + a[1]) + a[2]) + a[3];
      const double sB = ((bb[0] + bb[1]) + bb[2]) + bb[3];
      const double qA = sA * ee.x, qB = sB * ee.y;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const double fa = sA >= 4.0 ? 1.0 : (qA + a[q] * e1A);   // an all-missing subtree contributes exactly 1
        const double fb = sB >= 4.0 ? 1.0 : (qB + bb[q] * e1B);
        pv[q] = fa * fb;
      }
      double2* dst = reinterpret_cast<double2*>(clvLocus + ((size_t)ix.z * P + p) * 4);
      dst[0] = make_double2(pv[0], pv[1]);
      dst[1] = make_double2(pv[2], pv[3]);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; q++) sRoot[tid * 4 + q] = pv[q];
  __syncthreads();
  // root: as phase F of k_eval (.c:470-479)
  double term = 0.0;
  if (k > 0 && ph > 0) {
    double prob = 0.0;
    const int numConds = 4 * ph;
    for (int j = 0; j < numConds; j++) prob += sRoot[tid * 4 + j];
    term = log(prob / numConds) * cnt;
  }
  sTerm[tid] = term;
  __syncthreads();
  if (warp == 0 && lane < nl && mK[lane] > 0) {
    const double* tt = sTerm + (mColStart[lane] - b.firstCol);
    double lnl = 0.0;
    for (int j = 0; j < mP[lane]; j++) lnl += tt[j];
    d.lnL[b.firstLocus + lane] = lnl;
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
