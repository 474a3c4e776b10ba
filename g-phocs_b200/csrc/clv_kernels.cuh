// clv_kernels.cuh — data-likelihood kernels (Felsenstein pruning under JC69, phase-averaged root).
//
// Replaces computeLocusDataLikelihood + computeConditionalJC_new + computeSubtreeConditionals_new +
// computeEdgeConditionalJC (LocusDataLikelihood.c:426-483, 1559-1673, 1831-1848) for ALL loci in one launch.
//
// Layout in HBM:
//   columns   every live phased pattern of every locus is one "column"; loci own contiguous column ranges
//             colStart[l] .. colStart[l+1]
//   leaves    4-bit base masks (T=1,C=2,A=4,G=8,N=15), 16 leaves per 64-bit word, word-major
//             leafWords[w][column]  -> one coalesced 8-byte load per thread covers 16 leaves
//   internal  clv[colStart[l]*NI*8 + (((node-n)*2 + buf)*P_l + p)*4 + base]  (fp64) — for one
//             (locus,node,buffer) the P_l columns are contiguous 32-byte records, so a warp of
//             column-threads issues fully used 32-byte sectors whatever the tree shape
//   trees     father/left/right int16, age fp64, flag byte per node (tree_ops.cuh), locus-major
//
// Execution: one CTA of 128 threads owns a host-packed batch of whole loci whose columns fit the CTA
// (or one oversized locus, walked in 128-column chunks).  The CTA stages the batch's topology in
// shared memory, marks dirty nodes and their ancestors, orders them (post-order DFS restricted to the
// marked set), computes the JC69 edge probabilities of exactly those nodes once per locus into shared
// memory, then every thread walks its locus' schedule for its own column: children come from the
// leaf mask, from registers (the node computed just before), or from HBM; the new vector goes to the
// node's *other* buffer, so accept/reject never copies.  The root step sums 4*phases conditionals per
// phase group, takes count*log, and reduces per locus in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tree_ops.cuh"

namespace gphocs {

constexpr int kThreads = 128;      // threads (= columns) per CTA
constexpr int kMaxBatchLoci = 16;  // loci per CTA batch

struct Batch {
  int firstLocus, numLoci;
  int firstCol, numCols;  // numCols > kThreads only for a single oversized locus
  int scratchOff;         // offset (in columns) into rootScratch for oversized loci, else -1
  int pad;
};

struct StoreDev {
  int L, n, N, NI, W;
  long long Ct;
  const int* colStart;
  const unsigned long long* leafWords;
  const int* grpPhases;
  const int* grpCount;
  double* clv;
  int16_t *father, *left, *right, *svFather, *svLeft, *svRight;
  double *age, *svAge;
  uint8_t* flags;
  int *root, *savedRoot;
  double *rate, *lnL, *savedLnL;
  double* rootScratch;
  const uint8_t* active;  // per-locus mask or nullptr
  double* ctaSum;         // one partial sum per batch
};

struct SchedEntry {
  int16_t node, left, right;
  uint16_t info;  // bit0 dst buffer, bit1 left-child buffer, bit2 right-child buffer
  double e0L, e0R;
};

__host__ __device__ inline size_t evalSmemBytes(int n) {
  const int N = 2 * n - 1, NI = n - 1;
  size_t perLocus = (size_t)3 * N * sizeof(int16_t)  // father,left,right
                    + 2 * N                          // flags, need
                    + (size_t)(NI + 1) * sizeof(int16_t) * 2  // schedule node list + DFS stack
                    + (size_t)NI * sizeof(SchedEntry);
  perLocus = (perLocus + 15) & ~(size_t)15;
  return perLocus * kMaxBatchLoci + kMaxBatchLoci * 48 + (size_t)kThreads * 5 * sizeof(double) + 64;
}

__device__ __forceinline__ TreeView deviceView(const StoreDev& d, int l) {
  TreeView t;
  const size_t o = (size_t)l * d.N;
  t.father = d.father + o; t.left = d.left + o; t.right = d.right + o;
  t.svFather = d.svFather + o; t.svLeft = d.svLeft + o; t.svRight = d.svRight + o;
  t.age = d.age + o; t.svAge = d.svAge + o; t.flags = d.flags + o;
  t.root = d.root + l; t.savedRoot = d.savedRoot + l;
  t.lnL = d.lnL + l; t.savedLnL = d.savedLnL + l; t.rate = d.rate + l;
  t.numLeaves = d.n;
  t.numPatterns = d.colStart[l + 1] - d.colStart[l];
  return t;
}

// One thread per locus segment executes that locus' edit records in order (adjustGenNodeAge,
// executeGenSPR, scaleAllNodeAges, resetSaved, revertToSaved, setLocusMutationRate).
__global__ void k_apply_ops(StoreDev d, const Op* __restrict__ ops, const int* __restrict__ segStart, int nSegs,
                            int* __restrict__ status) {
  const int seg = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= nSegs) return;
  const int o0 = segStart[seg], o1 = segStart[seg + 1];
  const TreeView t = deviceView(d, ops[o0].locus);
  for (int o = o0; o < o1; o++) status[o] = applyOp(t, ops[o]);
}

// computeEdgeConditionalJC (.c:1831-1848): off-diagonal JC69 transition probability
__device__ __forceinline__ double edgeProb(double edgeLength) {
  if (edgeLength < 1e-100) return 0.0;
  return (1.0 - exp(-4.0 * edgeLength / 3.0)) / 4.0;
}

// computeSubtreeConditionals_new (.c:1650-1673)
__device__ __forceinline__ void foldChild(const double (&son)[4], double (&parent)[4], double e0) {
  const double s = ((son[0] + son[1]) + son[2]) + son[3];
  if (s >= 4.0) return;  // all-missing subtree
  const double e1 = 1.0 - 4.0 * e0;
  const double q = s * e0;
#pragma unroll
  for (int b = 0; b < 4; b++) parent[b] *= (q + son[b] * e1);
}

__device__ __forceinline__ void loadClv(const double* p, double (&v)[4]) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void storeClv(double* p, const double (&v)[4]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}

__global__ void __launch_bounds__(kThreads)
k_eval(StoreDev d, const Batch* __restrict__ batches, int batchBase, int useOld, int onlyLocus) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Batch b = batches[batchBase + blockIdx.x];
  const int tid = threadIdx.x;
  const int n = d.n, N = d.N, NI = d.NI, nl = b.numLoci;

  // ---- carve shared memory
  size_t perLocus = (size_t)3 * N * sizeof(int16_t) + 2 * N + (size_t)(NI + 1) * sizeof(int16_t) * 2 +
                    (size_t)NI * sizeof(SchedEntry);
  perLocus = (perLocus + 15) & ~(size_t)15;
  unsigned char* base = smem;
  double* sRoot = reinterpret_cast<double*>(base);              // [kThreads][4]
  double* sTerm = sRoot + kThreads * 4;                         // [kThreads]
  base += (size_t)kThreads * 5 * sizeof(double);
  int* mColStart = reinterpret_cast<int*>(base);                // per-slot metadata
  int* mP = mColStart + kMaxBatchLoci;
  int* mK = mP + kMaxBatchLoci;
  int* mRoot = mK + kMaxBatchLoci;
  int* mActive = mRoot + kMaxBatchLoci;
  double* mRate = reinterpret_cast<double*>(mActive + kMaxBatchLoci);
  double* mLnL = mRate + kMaxBatchLoci;
  base += kMaxBatchLoci * 48;
  auto slotBase = [&](int s) { return base + perLocus * s; };
  auto sSched = [&](int s) { return reinterpret_cast<SchedEntry*>(slotBase(s)); };
  auto sFather = [&](int s) { return reinterpret_cast<int16_t*>(slotBase(s) + (size_t)NI * sizeof(SchedEntry)); };
  auto sLeft = [&](int s) { return sFather(s) + N; };
  auto sRight = [&](int s) { return sFather(s) + 2 * N; };
  auto sList = [&](int s) { return sFather(s) + 3 * N; };            // [NI+1] post-order node list
  auto sStack = [&](int s) { return sFather(s) + 3 * N + NI + 1; };  // [NI+1]
  auto sFlags = [&](int s) { return reinterpret_cast<uint8_t*>(sFather(s) + 3 * N + 2 * (NI + 1)); };
  auto sNeed = [&](int s) { return sFlags(s) + N; };

  // ---- phase 0: per-locus metadata
  if (tid < nl) {
    const int l = b.firstLocus + tid;
    const int c0 = d.colStart[l];
    mColStart[tid] = c0;
    mP[tid] = d.colStart[l + 1] - c0;
    mRoot[tid] = d.root[l];
    mRate[tid] = d.rate[l];
    int act = (mP[tid] > 0) && (d.root[l] >= 0);
    if (d.active && !d.active[l]) act = 0;
    if (onlyLocus >= 0 && l != onlyLocus) act = 0;
    mActive[tid] = act;
    mK[tid] = 0;
    mLnL[tid] = 0.0;
  }
  // ---- phase 1: stage topology + flags (contiguous in HBM across the batch's loci -> coalesced)
  {
    const size_t g0 = (size_t)b.firstLocus * N;
    for (int i = tid; i < nl * N; i += kThreads) {
      const int s = i / N, v = i - s * N;
      sFather(s)[v] = d.father[g0 + i];
      sLeft(s)[v] = d.left[g0 + i];
      sRight(s)[v] = d.right[g0 + i];
      sFlags(s)[v] = d.flags[g0 + i];
      sNeed(s)[v] = 0;
    }
  }
  __syncthreads();
  // ---- phase 2: mark dirty nodes and their ancestors (computeConditionalJC_new's recursion condition, .c:1583)
  for (int i = tid; i < nl * N; i += kThreads) {
    const int s = i / N, v = i - s * N;
    if (!mActive[s]) continue;
    if (!useOld) {
      if (v >= n) sNeed(s)[v] = 1;
    } else if (sFlags(s)[v] & F_RECALC) {
      int u = v < n ? sFather(s)[v] : v;  // a moved leaf dirties its father (.c:1569-1575)
      uint8_t* need = sNeed(s);
      const int16_t* fa = sFather(s);
      while (u >= 0 && !need[u]) {
        need[u] = 1;
        u = fa[u];
      }
    }
  }
  __syncthreads();
  // ---- phase 3: per locus, post-order over the marked set; flip buffers of scheduled nodes
  if (tid < nl && mActive[tid]) {
    const int s = tid, l = b.firstLocus + s;
    uint8_t* need = sNeed(s);
    uint8_t* fl = sFlags(s);
    int16_t* list = sList(s);
    int16_t* stack = sStack(s);
    const int16_t *le = sLeft(s), *ri = sRight(s);
    int k = 0, sp = 0;
    const int root = mRoot[s];
    if (root >= n && need[root]) stack[sp++] = (int16_t)root;
    while (sp > 0) {
      const int v = stack[sp - 1];
      if (need[v] == 1) {
        need[v] = 2;
        const int r = ri[v], lf = le[v];
        if (r >= n && need[r] == 1) stack[sp++] = (int16_t)r;
        if (lf >= n && need[lf] == 1) stack[sp++] = (int16_t)lf;
      } else {
        sp--;
        list[k++] = (int16_t)v;
      }
    }
    uint8_t* gflags = d.flags + (size_t)l * N;
    for (int e = 0; e < k; e++) {
      const int v = list[e];
      uint8_t f = fl[v];
      if (!(f & F_RECALC)) {  // copyNodeConditionals: flip once per proposal
        f = (uint8_t)((f ^ F_SEL) | F_RECALC);
        fl[v] = f;
        gflags[v] = f;
      }
    }
    d.savedLnL[l] = d.lnL[l];  // always, even when nothing is recomputed (.c:440)
    mLnL[s] = d.lnL[l];
    mK[s] = k;
  }
  __syncthreads();
  // ---- phase 4: edge probabilities, once per (locus, scheduled node)
  for (int i = tid; i < nl * NI; i += kThreads) {
    const int s = i / NI, e = i - s * NI;
    if (e >= mK[s]) continue;
    const int v = sList(s)[e];
    const int lf = sLeft(s)[v], r = sRight(s)[v];
    const double* age = d.age + (size_t)(b.firstLocus + s) * N;
    const double av = age[v], rate = mRate[s];
    SchedEntry en;
    en.node = (int16_t)v; en.left = (int16_t)lf; en.right = (int16_t)r;
    const uint8_t* fl = sFlags(s);
    en.info = (uint16_t)((fl[v] & 1) | ((fl[lf] & 1) << 1) | ((fl[r] & 1) << 2));
    en.e0L = edgeProb(rate * (av - age[lf]));
    en.e0R = edgeProb(rate * (av - age[r]));
    sSched(s)[e] = en;
  }
  __syncthreads();

  // ---- phase 5: every thread walks its locus' schedule for its own column
  const int numChunks = (b.numCols + kThreads - 1) / kThreads;
  const bool oversized = b.scratchOff >= 0;
  for (int chunk = 0; chunk < numChunks; chunk++) {
    const int colInBatch = chunk * kThreads + tid;
    const bool live = colInBatch < b.numCols;
    const int c = b.firstCol + colInBatch;
    int s = 0;
    if (live) {
      while (s + 1 < nl && c >= mColStart[s + 1]) s++;
    }
    double pv[4] = {0.0, 0.0, 0.0, 0.0};
    int k = 0;
    if (live && mActive[s]) {
      k = mK[s];
      const int P = mP[s], p = c - mColStart[s];
      const unsigned long long w0 = d.leafWords[c];
      const unsigned long long w1 = d.W > 1 ? d.leafWords[(size_t)d.Ct + c] : 0ull;
      double* clvL = d.clv + (size_t)mColStart[s] * NI * 8;
      const SchedEntry* sched = sSched(s);
      int prevNode = -1;
      auto childValue = [&](int child, int buf, double (&v)[4]) {
        if (child < n) {
          const int w = child >> 4;
          const unsigned long long word = w == 0 ? w0 : (w == 1 ? w1 : d.leafWords[(size_t)w * d.Ct + c]);
          const unsigned code = (unsigned)(word >> ((child & 15) * 4)) & 15u;
#pragma unroll
          for (int q = 0; q < 4; q++) v[q] = (code >> q) & 1u ? 1.0 : 0.0;
        } else if (child == prevNode) {
#pragma unroll
          for (int q = 0; q < 4; q++) v[q] = pv[q];
        } else {
          loadClv(clvL + ((size_t)((child - n) * 2 + buf) * P + p) * 4, v);
        }
      };
      for (int e = 0; e < k; e++) {
        const SchedEntry en = sched[e];
        double a[4], bb[4], v[4] = {1.0, 1.0, 1.0, 1.0};
        childValue(en.left, (en.info >> 1) & 1, a);
        childValue(en.right, (en.info >> 2) & 1, bb);
        foldChild(a, v, en.e0L);
        foldChild(bb, v, en.e0R);
        storeClv(clvL + ((size_t)((en.node - n) * 2 + (en.info & 1)) * P + p) * 4, v);
        prevNode = en.node;
#pragma unroll
        for (int q = 0; q < 4; q++) pv[q] = v[q];
      }
    }
    // ---- phase 6: root conditionals -> shared (or scratch for an oversized locus)
    if (!oversized) {
#pragma unroll
      for (int q = 0; q < 4; q++) sRoot[tid * 4 + q] = pv[q];
    } else if (live && k > 0) {
      double* dst = d.rootScratch + ((size_t)b.scratchOff + colInBatch) * 4;
#pragma unroll
      for (int q = 0; q < 4; q++) dst[q] = pv[q];
    }
  }
  __syncthreads();

  if (!oversized) {
    // sum over the 4*phases root conditionals of each phase group, in the reference's order (.c:470-479)
    const bool live = tid < b.numCols;
    const int c = b.firstCol + tid;
    int s = 0;
    if (live) while (s + 1 < nl && c >= mColStart[s + 1]) s++;
    double term = 0.0;
    if (live && mActive[s] && mK[s] > 0) {
      const int ph = d.grpPhases[c];
      if (ph > 0) {
        double prob = 0.0;
        const int numConds = 4 * ph;
        for (int j = 0; j < numConds; j++) prob += sRoot[tid * 4 + j];
        term = log(prob / numConds) * d.grpCount[c];
      }
    }
    sTerm[tid] = term;
    __syncthreads();
    if (live && mActive[s] && mK[s] > 0 && c == mColStart[s]) {
      const int P = mP[s];
      const int* ph = d.grpPhases + c;
      double lnl = 0.0;
      for (int j = 0; j < P; j++)
        if (ph[j] > 0) lnl += sTerm[tid + j];
      d.lnL[b.firstLocus + s] = lnl;
      mLnL[s] = lnl;
    }
    __syncthreads();
    if (tid == 0) {
      double sum = 0.0;
      for (int j = 0; j < nl; j++)
        if (mActive[j]) sum += mLnL[j];
      d.ctaSum[batchBase + blockIdx.x] = sum;
    }
  } else {
    // oversized locus: groups may straddle chunks; reduce from scratch with a fixed-order block tree
    double acc = 0.0;
    const bool changed = mActive[0] && mK[0] > 0;
    if (changed) {
      const int P = mP[0], c0 = mColStart[0];
      const double* src = d.rootScratch + (size_t)b.scratchOff * 4;
      for (int p = tid; p < P; p += kThreads) {
        const int ph = d.grpPhases[c0 + p];
        if (ph > 0) {
          double prob = 0.0;
          const int numConds = 4 * ph;
          for (int j = 0; j < numConds; j++) prob += src[(size_t)p * 4 + j];
          acc += log(prob / numConds) * d.grpCount[c0 + p];
        }
      }
    }
    sTerm[tid] = acc;
    __syncthreads();
    for (int off = kThreads / 2; off > 0; off >>= 1) {
      if (tid < off) sTerm[tid] += sTerm[tid + off];
      __syncthreads();
    }
    if (tid == 0) {
      if (changed) {
        d.lnL[b.firstLocus] = sTerm[0];
        mLnL[0] = sTerm[0];
      }
      d.ctaSum[batchBase + blockIdx.x] = mActive[0] ? mLnL[0] : 0.0;
    }
  }
}

// fixed-order tree reduction of the per-CTA partial sums -> out[0]
__global__ void __launch_bounds__(1024) k_reduce_sum(const double* __restrict__ in, int n, double* __restrict__ out) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) acc += in[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 512; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

__global__ void k_set_mask(uint8_t* __restrict__ mask, const int* __restrict__ ids, int nIds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nIds) mask[ids[i]] = 1;
}

__global__ void k_gather_f64(const double* __restrict__ src, const int* __restrict__ ids, int nIds,
                             double* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nIds) dst[i] = src[ids[i]];
}

}  // namespace gphocs
