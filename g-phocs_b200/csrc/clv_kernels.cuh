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
//   trees     one 8-byte record {father,left,right:int16, flags:u8} + one fp64 age per node, locus-major
//
// Execution: one CTA of 128 threads owns a host-packed batch of whole loci whose columns fit the CTA
// (or one oversized locus, walked in 128-column chunks).
//   A  stage the batch's node records and ages in shared memory (coalesced 8-byte loads)
//   B  mark dirty nodes and their ancestors
//   C  flip the destination buffers of marked nodes; count marked nodes per subtree (shared atomics)
//   D  every marked node finds its own position in a post-order that visits the heavier child first by
//      walking up its ancestors (no serial traversal), and writes its schedule entry: destination,
//      where each child's vector comes from, JC69 edge terms
//   E  every thread walks its locus' schedule for its own column.  The vector computed last stays in
//      registers: in a post-order it is a child of the next entry unless that entry starts a new subtree.
//      Only a result whose sibling subtree is computed after it is parked on a per-column stack in shared
//      memory (depth <= log2(n) because the heavier child goes first), leaves are expanded from their
//      4-bit masks, clean children (incremental evaluation) come from HBM: in a full evaluation a child is
//      never re-read from HBM and about a third of the nodes touch shared memory at all.  The new vector
//      is written with one 32-byte store to the node's *other* buffer, so accept/reject never copies.
//   F  root: sum 4*phases conditionals per phase group, count*log, per-locus sum in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tree_ops.cuh"

namespace gphocs {

constexpr int kThreads = 128;      // threads (= columns) per CTA
constexpr int kWarps = kThreads / 32;
constexpr int kMaxBatchLoci = 16;  // loci per CTA batch
#ifndef GPHOCS_KSTACK
#define GPHOCS_KSTACK 3
#endif
#ifndef GPHOCS_EVAL_MINBLOCKS
#define GPHOCS_EVAL_MINBLOCKS 0
#endif
#if GPHOCS_EVAL_MINBLOCKS > 0
#define GPHOCS_EVAL_BOUNDS __launch_bounds__(kThreads, GPHOCS_EVAL_MINBLOCKS)
#else
#define GPHOCS_EVAL_BOUNDS __launch_bounds__(kThreads)
#endif
constexpr int kStack = GPHOCS_KSTACK;  // parked results per column in shared memory; deeper ones are re-read from HBM

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
  NodeRec *node, *saved;
  double *age, *svAge;
  int *root, *savedRoot;
  double *rate, *lnL, *savedLnL;
  double* rootScratch;
  const uint8_t* active;  // per-locus mask or nullptr
  double* ctaSum;         // one partial sum per batch
  // optional accounting for the roofline of incremental evaluations (nullptr: off): [0] += loci re-evaluated,
  // [1] += their algorithmic bytes 32*P*(2k+1), k = conditional vectors recomputed
  unsigned long long* evalCounters;
};

// One schedule entry = one node to (re)compute; entries are ordered so that children precede parents.
enum : uint32_t { SRC_LEAF = 0, SRC_STACK = 1, SRC_GLOBAL = 2, SRC_TOP = 3 };
constexpr uint32_t kRow = kThreads * 16;       // bytes per stack row (one double2 per column)
constexpr uint32_t kHi = kStack * kRow;        // distance from the lo half of a vector to its hi half
struct __align__(16) SchedEntry {
  double e0A, e1A, e0B, e1B;  // JC69 edge terms of the two children: p and 1-4p (.c:1596-1602)
  // where a child's vector comes from, ready to use by the column threads:
  //   leaf    byte offset of its 32-bit code word row | bit shift << 16
  //   stack   (slot * row bytes) << 16
  //   global  byte offset of the source record from the column's base
  //   top     the vector of the previous entry, still in registers
  uint32_t offA, offB;
  uint32_t dstOff;            // byte offset of the destination record from the column's base
  uint32_t ctl;               // bits 0-1 kind of A, 2-3 kind of B, bits 16-31 stack byte offset to park the result at (0xffff: none)
};
static_assert(sizeof(SchedEntry) == 48, "schedule entry layout");
// the same without the two 1-4p terms, which the walking thread then derives from p (same expression, same bits): the
// fused sweep kernel keeps more CTAs per SM with 32-byte entries
struct __align__(16) SchedEntryCompact {
  double e0A, e0B;
  uint32_t offA, offB, dstOff, ctl;
};
static_assert(sizeof(SchedEntryCompact) == 32, "compact schedule entry layout");

struct EvalSmem {
  size_t perScratch, offAge, offNode, offSize, offWalk, offNeed;  // per-locus scheduling scratch
  size_t perSched;                                                // per-locus schedule
  size_t offStack, offSched, offWords, offTerm, offMeta, offList, total;   // CTA regions
  int W32;
};
// Shared-memory plan for a CTA that holds up to `maxLoci` loci of `n` leaves.
__host__ __device__ inline EvalSmem evalSmemLayout(int n, int maxLoci) {
  const int N = 2 * n - 1, NI = n - 1;
  EvalSmem m;
  m.W32 = (n + 7) / 8;
  m.offAge = 0;                                            // [N] double
  m.offNode = m.offAge + (size_t)N * sizeof(double);       // [N] NodeRec
  m.offSize = m.offNode + (size_t)N * sizeof(NodeRec);     // [NI] int: marked nodes per subtree
  m.offWalk = m.offSize + (size_t)NI * sizeof(int);        // [N] uint32: father+1 | contribution << 16
  m.offNeed = m.offWalk + (size_t)N * sizeof(uint32_t);    // [N] uint8
  m.perScratch = (m.offNeed + (size_t)N + 15) & ~(size_t)15;
  m.perSched = (size_t)NI * sizeof(SchedEntry);
  // Region 0 is used twice: first as the scheduling scratch of the batch's loci, then (phase E) as the
  // per-column stack: lo halves [kStack][kThreads] double2, then the hi halves.
  // The root vectors reuse its first rows after the walk.
  const size_t stackBytes = 2 * (size_t)kHi, scratchBytes = m.perScratch * (size_t)maxLoci;
  m.offStack = 0;
  m.offSched = stackBytes > scratchBytes ? stackBytes : scratchBytes;
  m.offWords = m.offSched + m.perSched * (size_t)maxLoci;           // [W32][kThreads] uint32 leaf codes
  m.offTerm = m.offWords;                                           // [kThreads] double, after the walk
  const size_t wordBytes = (size_t)m.W32 * kThreads * 4, termBytes = (size_t)kThreads * 8;
  m.offMeta = m.offWords + (wordBytes > termBytes ? wordBytes : termBytes);
  m.offList = m.offMeta + (size_t)kMaxBatchLoci * 48;              // [1 + maxLoci*NI] uint32: marked nodes of the batch
  m.total = m.offList + (1 + (size_t)maxLoci * NI) * sizeof(uint32_t);
  m.total = (m.total + 15) & ~(size_t)15;
  return m;
}
__host__ __device__ inline size_t evalSmemBytes(int n, int maxLoci = kMaxBatchLoci) { return evalSmemLayout(n, maxLoci).total; }

__device__ __forceinline__ TreeView deviceView(const StoreDev& d, int l) {
  TreeView t;
  const size_t o = (size_t)l * d.N;
  t.node = d.node + o; t.saved = d.saved + o;
  t.age = d.age + o; t.svAge = d.svAge + o;
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

// The same for records that arrive sorted by locus (gphocsStoreApplyOpsAsync): the thread that holds the first record
// of a locus replays that locus' records; no segment table crosses PCIe.  *bad counts records outside the store or tree.
__global__ void k_apply_ops_sorted(StoreDev d, const Op* __restrict__ ops, int nOps, int* __restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nOps) return;
  const int locus = ops[i].locus;
  if (i > 0 && ops[i - 1].locus == locus) return;
  if (locus < 0 || locus >= d.L || (i > 0 && ops[i - 1].locus > locus)) { atomicAdd(bad, 1); return; }
  const TreeView t = deviceView(d, locus);
  for (int o = i; o < nOps && ops[o].locus == locus; o++) {
    const Op op = ops[o];
    bool ok = op.type >= OP_ADJUST_AGE && op.type <= OP_SET_RATE;
    if (op.type == OP_ADJUST_AGE) ok = op.a >= 0 && op.a < d.N;
    if (op.type == OP_SPR) ok = op.a >= 0 && op.a < d.N && op.b >= 0 && op.b < d.N && op.a != op.b && t.node[op.a].father >= 0;
    if (!ok) { atomicAdd(bad, 1); return; }
    applyOp(t, op);
  }
}

// genealogies host -> device: topology fields, ages and roots of the listed loci; flag bytes (buffer
// selectors, dirty marks) stay as they are on the device
__global__ void k_set_trees(StoreDev d, const int* __restrict__ ids, const int16_t* __restrict__ topo,
                            const double* __restrict__ age, const int* __restrict__ root, int nLoci) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)nLoci * d.N) return;
  const int k = (int)(i / d.N), v = (int)(i - (size_t)k * d.N);
  const int l = ids[k];
  const size_t o = (size_t)l * d.N + v;
  NodeRec r = d.node[o];
  r.father = topo[i * 3]; r.left = topo[i * 3 + 1]; r.right = topo[i * 3 + 2];
  d.node[o] = r;
  d.age[o] = age[i];
  if (v == 0) d.root[l] = root[k];
}

// the same for loci 0..nLoci-1 whose int32 topology arrays were copied to the device as they are (ages and roots
// went straight to their final arrays)
__global__ void k_set_topology32(StoreDev d, const int* __restrict__ father, const int* __restrict__ left,
                                 const int* __restrict__ right, size_t count, int* __restrict__ bad) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int f = father[i], l = left[i], rr = right[i];
  if (f < -1 || f >= d.N || l < -1 || l >= d.N || rr < -1 || rr >= d.N) { atomicAdd(bad, 1); return; }
  NodeRec r = d.node[i];
  r.father = (int16_t)father[i]; r.left = (int16_t)left[i]; r.right = (int16_t)right[i];
  d.node[i] = r;
}

// the same for 16-bit (father, left, right) triples; *bad counts ids outside [-1, N-1]
__global__ void k_set_topology16(StoreDev d, const int16_t* __restrict__ topo, size_t count, int* __restrict__ bad) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int16_t f = topo[3 * i], l = topo[3 * i + 1], r = topo[3 * i + 2];
  if (f < -1 || f >= d.N || l < -1 || l >= d.N || r < -1 || r >= d.N) { atomicAdd(bad, 1); return; }
  NodeRec rec = d.node[i];
  rec.father = f; rec.left = l; rec.right = r;
  d.node[i] = rec;
}

// computeEdgeConditionalJC (.c:1831-1848): off-diagonal JC69 transition probability
__device__ __forceinline__ double edgeProb(double edgeLength) {
  if (edgeLength < 1e-100) return 0.0;
  return (1.0 - exp(-4.0 * edgeLength / 3.0)) / 4.0;
}

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 ldsD2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void stsD2(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}

__device__ __forceinline__ void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint32_t ldsU32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void stsU32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ldsU4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
// one (node, column) record = 4 doubles = one 32-byte sector: a single 256-bit access per thread (sm_100)
__device__ __forceinline__ void ldgD4(const void* p, double (&v)[4]) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void stgD4(void* p, const double (&v)[4]) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
// rare path of the column walk
__device__ __forceinline__ void missingSubtree(const double (&a)[4], const double (&b)[4], double sA, double sB, double qA,
                                            double qB, double e1A, double e1B, double (&v)[4]) {
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const double fa = sA >= 4.0 ? 1.0 : (qA + a[q] * e1A);
    const double fb = sB >= 4.0 ? 1.0 : (qB + b[q] * e1B);
    v[q] = fa * fb;
  }
}

// Phase E of k_eval, shared with the fused sweep kernel (sweep_kernels.cuh): one thread walks the k schedule entries at
// shared-memory address `entry` for its own column.  clvCol = the column's base in the conditional-vector store,
// myStack / myWords = this thread's column of the parking stack and of the leaf-code rows.  pv returns the vector
// computed last (the root's, when the schedule ends at the root).  HI = bytes from the lo half of a parked vector to
// its hi half (= rows of the stack * kRow).
template <uint32_t HI = kHi, bool COMPACT = false>
__device__ __forceinline__ void columnWalk(uint32_t entry, int k, char* clvCol, uint32_t myStack, uint32_t myWords, bool useOld,
                                           double (&pv)[4]) {
  constexpr uint32_t kEntryBytes = COMPACT ? sizeof(SchedEntryCompact) : sizeof(SchedEntry);
  // pv = the vector computed by the previous entry (the register-resident top of the stack)
  auto childValue = [&](uint32_t kind, uint32_t off, double (&v)[4]) {
    if (kind == SRC_TOP) {
#pragma unroll
      for (int q = 0; q < 4; q++) v[q] = pv[q];
    } else if (kind == SRC_LEAF) {
      // conditional vector of a leaf from its 4-bit base mask (computeLeafConditionals, .c:1336-1386):
      // 1.0 = 0x3ff00000'00000000 where the bit is set, 0.0 elsewhere
      const uint32_t m = ldsU32(myWords + (off & 0xffffu)) >> (off >> 16);
#pragma unroll
      for (int q = 0; q < 4; q++) v[q] = __hiloint2double((int)(((m >> q) & 1u) * 0x3ff00000u), 0);
    } else if (kind == SRC_STACK) {
      const uint32_t a = myStack + (off >> 16);
      const double2 x = ldsD2(a), y = ldsD2(a + HI);
      v[0] = x.x; v[1] = x.y; v[2] = y.x; v[3] = y.y;
    } else {
      ldgD4(clvCol + off, v);
    }
  };
  if (useOld) {   // clean children are the only HBM reads of a proposal: request them all before the first use
    for (int e = 0; e < k; e++) {
      const uint4 ix = ldsU4(entry + e * kEntryBytes + kEntryBytes - 16);
      if ((ix.w & 3u) == SRC_GLOBAL) prefetchL2(clvCol + ix.x);
      if (((ix.w >> 2) & 3u) == SRC_GLOBAL) prefetchL2(clvCol + ix.y);
    }
  }
  for (int e = 0; e < k; e++, entry += kEntryBytes) {
    double2 eA, eB;                                            // (e0A,e1A), (e0B,e1B)
    if (COMPACT) {
      const double2 e0 = ldsD2(entry);
      eA.x = e0.x; eA.y = 1.0 - 4.0 * e0.x;
      eB.x = e0.y; eB.y = 1.0 - 4.0 * e0.y;
    } else {
      eA = ldsD2(entry); eB = ldsD2(entry + 16);
    }
    const uint4 ix = ldsU4(entry + kEntryBytes - 16);          // offA, offB, dstOff, ctl
    double a[4], bb[4], v[4];
    childValue(ix.w & 3u, ix.x, a);
    childValue((ix.w >> 2) & 3u, ix.y, bb);
    // computeSubtreeConditionals_new (.c:1650-1673) for both children
    const double sA = ((a[0] + a[1]) + a[2]) + a[3];
    const double sB = ((bb[0] + bb[1]) + bb[2]) + bb[3];
    const double qA = sA * eA.x, qB = sB * eB.x;
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] = (qA + a[q] * eA.y) * (qB + bb[q] * eB.y);
    // sums are positive, so s >= 4.0 can be read off the high word (4.0 = 0x40100000'00000000)
    if (__builtin_expect((__double2hiint(sA) >= 0x40100000) | (__double2hiint(sB) >= 0x40100000), 0)) {
      // a child whose conditionals sum to 4 is an all-missing subtree and contributes exactly 1 (.c:1660-1663)
      missingSubtree(a, bb, sA, sB, qA, qB, eA.y, eB.y, v);
    }
    stgD4(clvCol + ix.z, v);
    const uint32_t push = ix.w >> 16;
    if (push != 0xffffu) {
      stsD2(myStack + push, v[0], v[1]);
      stsD2(myStack + push + HI, v[2], v[3]);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) pv[q] = v[q];
  }
}

__global__ void GPHOCS_EVAL_BOUNDS
k_eval(StoreDev d, const Batch* __restrict__ batches, int batchBase, int useOld, int onlyLocus, int maxLoci,
       int prefetchAhead) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Batch b = batches[batchBase + blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = d.n, N = d.N, NI = d.NI, nl = b.numLoci;
  const EvalSmem lay = evalSmemLayout(n, maxLoci);

  // ---- carve shared memory
  const uint32_t stackBase = smemAddr(smem + lay.offStack);
  double* sRoot = reinterpret_cast<double*>(smem + lay.offStack);        // [kThreads][4], after the walk
  double* sTerm = reinterpret_cast<double*>(smem + lay.offTerm);         // [kThreads]
  int* mColStart = reinterpret_cast<int*>(smem + lay.offMeta);           // per-slot metadata
  int* mP = mColStart + kMaxBatchLoci;
  int* mK = mP + kMaxBatchLoci;
  int* mRoot = mK + kMaxBatchLoci;
  int* mActive = mRoot + kMaxBatchLoci;
  double* mRate = reinterpret_cast<double*>(mActive + kMaxBatchLoci);
  double* mLnL = mRate + kMaxBatchLoci;
  int* sListCount = reinterpret_cast<int*>(smem + lay.offList);
  uint32_t* sList = reinterpret_cast<uint32_t*>(smem + lay.offList) + 1;
  auto sSched = [&](int s) { return reinterpret_cast<SchedEntry*>(smem + lay.offSched + lay.perSched * s); };
  auto sAge = [&](int s) { return reinterpret_cast<double*>(smem + lay.perScratch * s + lay.offAge); };
  auto sNode = [&](int s) { return reinterpret_cast<NodeRec*>(smem + lay.perScratch * s + lay.offNode); };
  auto sSize = [&](int s) { return reinterpret_cast<int*>(smem + lay.perScratch * s + lay.offSize); };
  auto sWalk = [&](int s) { return reinterpret_cast<uint32_t*>(smem + lay.perScratch * s + lay.offWalk); };
  auto sNeed = [&](int s) { return reinterpret_cast<uint8_t*>(smem + lay.perScratch * s + lay.offNeed); };

  // ---- early loads: what the column walk and the root step need from HBM depends only on the batch
  //      descriptor, so it is requested now and arrives while the schedule is being built
  const bool oversized = b.scratchOff >= 0;
  unsigned long long w0 = 0ull, w1 = 0ull;
  int ph = 0, cnt = 0;
  if (tid < b.numCols) {
    const int c = b.firstCol + tid;
    w0 = d.leafWords[c];
    if (d.W > 1) w1 = d.leafWords[(size_t)d.Ct + c];
    if (!oversized) { ph = d.grpPhases[c]; cnt = d.grpCount[c]; }
  }
  // ---- phase A: per-locus metadata; node records and ages staged with coalesced loads
  if (tid < nl) {
    const int l = b.firstLocus + tid;
    const int c0 = d.colStart[l];
    mColStart[tid] = c0;
    mP[tid] = d.colStart[l + 1] - c0;
    const int root = d.root[l];
    mRoot[tid] = root;
    mRate[tid] = d.rate[l];
    int act = (mP[tid] > 0) && (root >= n);
    if (d.active && !d.active[l]) act = 0;
    if (onlyLocus >= 0 && l != onlyLocus) act = 0;
    mActive[tid] = act;
    mK[tid] = 0;
    const double cur = d.lnL[l];
    mLnL[tid] = cur;
    if (act) d.savedLnL[l] = cur;  // always, even when nothing is recomputed (.c:440)
  }
  for (int s = warp; s < nl; s += kWarps) {
    const size_t g0 = (size_t)(b.firstLocus + s) * N;
    NodeRec* nd = sNode(s);
    double* age = sAge(s);
    uint8_t* need = sNeed(s);
    int* size = sSize(s);
    for (int v = lane; v < N; v += 32) {
      nd[v] = d.node[g0 + v];
      age[v] = d.age[g0 + v];
      need[v] = 0;
      if (v < NI) size[v] = 0;
    }
  }
  __syncthreads();
  // Pull the inputs of the batch that the CTA taking this one's place will own into L2 (128-byte lines,
  // spread over the CTA), so that CTA does not start with a chain of HBM round trips.
  if (prefetchAhead > 0 && onlyLocus < 0 && (int)blockIdx.x + prefetchAhead < (int)gridDim.x) {
    const Batch ahead = batches[batchBase + blockIdx.x + prefetchAhead];   // in L2: prefetched by an earlier CTA
    if (tid == 0) prefetchL2(batches + batchBase + blockIdx.x + 2 * prefetchAhead);
    const char* p0 = reinterpret_cast<const char*>(d.node + (size_t)ahead.firstLocus * N);
    const char* p1 = reinterpret_cast<const char*>(d.age + (size_t)ahead.firstLocus * N);
    const int recBytes = ahead.numLoci * N * 8;
    for (int o = tid * 128; o < recBytes; o += kThreads * 128) { prefetchL2(p0 + o); prefetchL2(p1 + o); }
    const int cols = min(ahead.numCols, kThreads);
    if (tid * 16 < cols) {   // 16 columns of 8 bytes per line
      for (int w = 0; w < d.W; w++) prefetchL2(d.leafWords + (size_t)w * d.Ct + ahead.firstCol + tid * 16);
    }
    if (tid * 32 < cols) {   // 32 columns of 4 bytes per line
      prefetchL2(d.grpPhases + ahead.firstCol + tid * 32);
      prefetchL2(d.grpCount + ahead.firstCol + tid * 32);
    }
    if (tid == 0) {
      prefetchL2(d.colStart + ahead.firstLocus); prefetchL2(d.root + ahead.firstLocus);
      prefetchL2(d.rate + ahead.firstLocus); prefetchL2(d.lnL + ahead.firstLocus);
    }
  }
  // ---- phase B: mark dirty nodes and their ancestors (computeConditionalJC_new's recursion condition, .c:1583)
  for (int s = warp; s < nl; s += kWarps) {
    if (!mActive[s]) continue;
    const NodeRec* nd = sNode(s);
    uint8_t* need = sNeed(s);
    for (int v = lane; v < N; v += 32) {
      if (!useOld) {
        if (v >= n) need[v] = 1;
      } else if (nd[v].flags & F_RECALC) {
        int u = v < n ? nd[v].father : v;  // a moved leaf dirties its father (.c:1569-1575)
        // lanes may walk the same ancestors at once: every writer stores the same 1 and a reader that misses it only
        // walks a little further (racecheck reports these read/write pairs as warnings; they are the only ones)
        for (int it = 0; u >= 0 && !need[u] && it < N; it++) {
          need[u] = 1;
          u = nd[u].father;
        }
      }
    }
  }
  if (tid == 0) *sListCount = 0;
  __syncthreads();
  // ---- phase C0: the marked nodes of the whole batch, compacted into one list so that every later phase keeps
  //      all 128 threads busy whether 4 nodes per locus are marked (a proposal) or all of them (full evaluation);
  //      flip the destination buffers of the marked nodes (copyNodeConditionals: once per proposal)
  for (int s = warp; s < nl; s += kWarps) {
    if (!mActive[s]) continue;
    NodeRec* nd = sNode(s);
    const uint8_t* need = sNeed(s);
    for (int v0 = n; v0 < N; v0 += 32) {
      const int v = v0 + lane;
      const bool marked = v < N && need[v];
      const unsigned ballot = __ballot_sync(0xffffffffu, marked);
      int base = 0;
      if (lane == 0 && ballot) base = atomicAdd(sListCount, __popc(ballot));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (marked) {
        sList[base + __popc(ballot & ((1u << lane) - 1u))] = (uint32_t)(s << 16 | v);
        uint8_t f = nd[v].flags;
        if (!(f & F_RECALC)) {
          f = (uint8_t)((f ^ F_SEL) | F_RECALC);
          nd[v].flags = f;
          d.node[(size_t)(b.firstLocus + s) * N + v].flags = f;
        }
      }
    }
  }
  __syncthreads();
  const int listCount = *sListCount;
  // ---- phase C: count the marked nodes of every subtree
  for (int j = tid; j < listCount; j += kThreads) {
    const int s = sList[j] >> 16, v = sList[j] & 0xffff;
    const NodeRec* nd = sNode(s);
    int* size = sSize(s);
    int a = v;
    for (int it = 0; a >= 0 && it < N; it++) {
      atomicAdd(&size[a - n], 1);
      a = nd[a].father;
    }
  }
  __syncthreads();
  // ---- phase D1: what a node contributes to the post-order start of everything below it.  The heavier
  //      child of a node is visited first; a node visited second starts after its sibling's subtree.
  for (int j = tid; j < listCount; j += kThreads) {
    const int s = sList[j] >> 16, v = sList[j] & 0xffff;
    const NodeRec* nd = sNode(s);
    const uint8_t* need = sNeed(s);
    const int* size = sSize(s);
    const int a = nd[v].father;
    uint32_t contrib = 0;
    if (a >= 0) {
      const int l = nd[a].left, r = nd[a].right;
      const int wl = (l >= n && need[l]) ? size[l - n] : 0;
      const int wr = (r >= n && need[r]) ? size[r - n] : 0;
      const int first = wl >= wr ? l : r;
      if (v != first) contrib = (uint32_t)(v == l ? wr : wl);
    }
    sWalk(s)[v] = (uint32_t)(a + 1) | (contrib << 16);
  }
  __syncthreads();
  // ---- phase D2: position and stack depth of every marked node, sources of its children, JC69 edge terms
  for (int j = tid; j < listCount; j += kThreads) {
    const int s = sList[j] >> 16, v = sList[j] & 0xffff;
    const NodeRec* nd = sNode(s);
    const uint8_t* need = sNeed(s);
    const int* size = sSize(s);
    const uint32_t* walk = sWalk(s);
    const double* age = sAge(s);
    const double rate = mRate[s];
    int start = 0, depth = 0;  // depth = results parked on the stack when v's subtree is entered
    {
      uint32_t w = walk[v];
      for (int it = 0; it < N; it++) {
        const uint32_t c = w >> 16;
        start += c;
        depth += c != 0;
        const int a = (int)(w & 0xffffu) - 1;
        if (a < 0) break;
        w = walk[a];
      }
    }
    const int l = nd[v].left, r = nd[v].right;
    const int wl = (l >= n && need[l]) ? size[l - n] : 0;
    const int wr = (r >= n && need[r]) ? size[r - n] : 0;
    const bool leftFirst = wl >= wr;
    const int A = leftFirst ? l : r, B = leftFirst ? r : l;  // visiting order
    const int wA = leftFirst ? wl : wr, wB = leftFirst ? wr : wl;
    const uint32_t strideBytes = (uint32_t)mP[s] * 32u;  // one (node, buffer) record of this locus
    auto record = [&](int x) { return (uint32_t)((x - n) * 2 + (nd[x].flags & F_SEL)) * strideBytes; };
    auto leafRef = [&](int x) { return (uint32_t)(x >> 3) * (kThreads * 4u) | ((uint32_t)(x & 7) * 4u) << 16; };
    // The entry right before v's is the root of the subtree visited last, so that child's vector is still in the
    // walking thread's registers.  When both subtrees are computed, A's result was parked on the stack at the depth
    // v's subtree was entered with (or, beyond kStack, is re-read from the record the thread has just written).
    uint32_t kindA, kindB, offA = 0, offB = 0;
    if (A < n) { kindA = SRC_LEAF; offA = leafRef(A); }
    else if (wA > 0 && wB == 0) { kindA = SRC_TOP; }
    else if (wA > 0 && depth < kStack) { kindA = SRC_STACK; offA = ((uint32_t)depth * kRow) << 16; }
    else { kindA = SRC_GLOBAL; offA = record(A); }  // clean child, or a result the thread parked in HBM
    if (B < n) { kindB = SRC_LEAF; offB = leafRef(B); }
    else if (wB > 0) { kindB = SRC_TOP; }
    else { kindB = SRC_GLOBAL; offB = record(B); }
    SchedEntry en;
    const double av = age[v];
    en.e0A = edgeProb(rate * (av - age[A]));
    en.e1A = 1.0 - 4.0 * en.e0A;
    en.e0B = edgeProb(rate * (av - age[B]));
    en.e1B = 1.0 - 4.0 * en.e0B;
    en.offA = offA; en.offB = offB;
    en.dstOff = record(v);
    // v's own result is parked iff v is visited first and its sibling's subtree is computed too
    uint32_t push = 0xffffu;
    {
      const int f = nd[v].father;
      if (f >= 0 && v != mRoot[s] && depth < kStack) {
        const int fl = nd[f].left, fr = nd[f].right;
        const int wfl = (fl >= n && need[fl]) ? size[fl - n] : 0;
        const int wfr = (fr >= n && need[fr]) ? size[fr - n] : 0;
        const int first = wfl >= wfr ? fl : fr;
        const int wSibling = v == fl ? wfr : wfl;
        if (v == first && wSibling > 0) push = (uint32_t)depth * kRow;
      }
    }
    en.ctl = kindA | (kindB << 2) | (push << 16);
    sSched(s)[start + size[v - n] - 1] = en;
    if (v == mRoot[s]) mK[s] = size[v - n];
  }
  __syncthreads();

  // ---- phase E: every thread walks its locus' schedule for its own column.  The scheduling scratch is dead;
  //      its space becomes the per-column stack.
  const int numChunks = (b.numCols + kThreads - 1) / kThreads;
  const uint32_t myStack = stackBase + tid * 16;
  const uint32_t myWords = smemAddr(smem + lay.offWords) + tid * 4;
  int s = 0, k = 0;
  for (int chunk = 0; chunk < numChunks; chunk++) {
    const int colInBatch = chunk * kThreads + tid;
    const bool live = colInBatch < b.numCols;
    const int c = b.firstCol + colInBatch;
    s = 0;
    if (live) {
      while (s + 1 < nl && c >= mColStart[s + 1]) s++;
    }
    double pv[4] = {0.0, 0.0, 0.0, 0.0};
    k = (live && mActive[s]) ? mK[s] : 0;
    if (chunk > 0 && live) {
      w0 = d.leafWords[c];
      w1 = d.W > 1 ? d.leafWords[(size_t)d.Ct + c] : 0ull;
    }
    if (k > 0) {
      const int p = c - mColStart[s];
      // this column's leaf masks, 8 leaves per 32-bit word
      stsU32(myWords, (uint32_t)w0);
      if (lay.W32 > 1) stsU32(myWords + kThreads * 4, (uint32_t)(w0 >> 32));
      if (lay.W32 > 2) stsU32(myWords + 2 * kThreads * 4, (uint32_t)w1);
      if (lay.W32 > 3) stsU32(myWords + 3 * kThreads * 4, (uint32_t)(w1 >> 32));
      for (int w = 2; w < d.W; w++) {
        const unsigned long long word = d.leafWords[(size_t)w * d.Ct + c];
        stsU32(myWords + (2 * w) * kThreads * 4, (uint32_t)word);
        if (2 * w + 1 < lay.W32) stsU32(myWords + (2 * w + 1) * kThreads * 4, (uint32_t)(word >> 32));
      }
      char* clvCol = reinterpret_cast<char*>(d.clv + (size_t)mColStart[s] * NI * 8 + (size_t)p * 4);
      columnWalk(smemAddr(sSched(s)), k, clvCol, myStack, myWords, useOld != 0, pv);
    }
    // ---- root conditionals -> shared (or scratch for an oversized locus)
    if (!oversized) {
      __syncthreads();  // the stack is dead from here on; its space takes the root vectors
#pragma unroll
      for (int q = 0; q < 4; q++) sRoot[tid * 4 + q] = pv[q];
    } else if (k > 0) {
      double* dst = d.rootScratch + ((size_t)b.scratchOff + colInBatch) * 4;
#pragma unroll
      for (int q = 0; q < 4; q++) dst[q] = pv[q];
    }
  }
  __syncthreads();

  // ---- phase F
  if (!oversized) {
    // sum over the 4*phases root conditionals of each phase group, in the reference's order (.c:470-479)
    double term = 0.0;
    if (k > 0 && ph > 0) {
      double prob = 0.0;
      const int numConds = 4 * ph;
      for (int j = 0; j < numConds; j++) prob += sRoot[tid * 4 + j];
      term = log(prob / numConds) * cnt;
    }
    sTerm[tid] = term;  // 0.0 for the other members of a phase group: adding it below is exact
    __syncthreads();
    if (warp == 0) {
      // per-locus sums in pattern order, then the CTA's partial sum in locus order (both fixed, deterministic)
      double lnl = 0.0;
      const bool mine = lane < nl && mActive[lane];
      if (mine) {
        lnl = mLnL[lane];
        if (mK[lane] > 0) {
          const int P = mP[lane];
          const double* t = sTerm + (mColStart[lane] - b.firstCol);
          lnl = 0.0;
          for (int j = 0; j < P; j++) lnl += t[j];
          d.lnL[b.firstLocus + lane] = lnl;
        }
      }
      double sum = 0.0;
      for (int j = 0; j < nl; j++) {
        const double x = __shfl_sync(0xffffffffu, lnl, j);
        const int on = __shfl_sync(0xffffffffu, (int)mine, j);
        if (on) sum += x;
      }
      if (lane == 0) d.ctaSum[batchBase + blockIdx.x] = sum;
      if (d.evalCounters && useOld) {
        unsigned long long by = mine && mK[lane] > 0 ? 32ull * (unsigned long long)mP[lane] * (2ull * (unsigned long long)mK[lane] + 1ull) : 0ull;
        unsigned cntEv = __popc(__ballot_sync(0xffffffffu, by != 0ull));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) by += __shfl_xor_sync(0xffffffffu, by, off);
        if (lane == 0 && cntEv) { atomicAdd(d.evalCounters, (unsigned long long)cntEv); atomicAdd(d.evalCounters + 1, by); }
      }
    }
  } else {
    // oversized locus: groups may straddle chunks; reduce from scratch with a fixed-order block tree
    double acc = 0.0;
    const bool changed = mActive[0] && mK[0] > 0;
    if (changed) {
      const int P = mP[0], c0 = mColStart[0];
      const double* src = d.rootScratch + (size_t)b.scratchOff * 4;
      for (int p = tid; p < P; p += kThreads) {
        const int phs = d.grpPhases[c0 + p];
        if (phs > 0) {
          double prob = 0.0;
          const int numConds = 4 * phs;
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
      if (d.evalCounters && useOld && changed) {   // the roofline accounting of the sampler (as for the loci of a shared batch above)
        atomicAdd(d.evalCounters, 1ull);
        atomicAdd(d.evalCounters + 1, 32ull * (unsigned long long)mP[0] * (2ull * (unsigned long long)mK[0] + 1ull));
      }
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
