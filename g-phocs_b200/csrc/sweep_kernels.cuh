// sweep_kernels.cuh — the two per-locus sweeps of an MCMC iteration in ONE launch, loci batched across a CTA.
//
//   UpdateGB_InternalNode  GPhoCS.c:2287-2424   one coalescence-time proposal per internal node
//   UpdateGB_MigSPR        GPhoCS.c:2598-2760   one prune-and-regraft proposal per node (models without migration bands)
//
// Loci are independent inside these sweeps (GPhoCS.c:2297, 2608: `omp parallel for` over loci), so a CTA keeps its
// batch of loci — the same batches k_eval works on: whole loci whose pattern columns fit the CTA's 128 threads — for
// the WHOLE of both sweeps.  Genealogies, population assignments, coal statistics and the leaf codes of the batch are
// staged once in shared memory; per step the CTA runs
//   team phase   8 threads per locus: accept / reject of the previous proposal (smpResolve), the next proposal
//                (smpAgeProposeBody / smpSprProposeBody) edited through tree_ops.cuh on the shared-memory genealogy,
//                dirty marking and compaction of the nodes to recompute (k_eval phases B, C0)
//   list phases  one thread per marked node of the batch: subtree sizes, post-order positions, schedule entries
//                with JC69 edge terms (k_eval phases C, D)
//   column phase one thread per pattern column walks its locus' schedule (columnWalk, shared with k_eval); the
//                conditional vectors stay in HBM/L2 — a CTA re-reads what it wrote a few microseconds earlier
//   root phase   phase-averaged root sums and per-locus log-likelihoods (k_eval phase F)
// and nothing but the final state goes back to HBM.  One launch replaces 2(n-1) + 2(2n-1) + 1 launches of the
// stepwise route (k_smp_age_propose / k_smp_spr_propose + k_eval per node), whose warp-per-locus proposal kernels left
// half of their lanes idle and re-staged every genealogy twice per step.
//
// Same chain as the stepwise route, bit for bit: random streams are keyed by (locus, step) with the stepwise route's
// step numbers; the column arithmetic is the same code; the coal statistic of a moved node's population is summed in
// the order of the stepwise route's 32-lane shuffle tree (teamPopStat).  tests/test_gpu_sampler.py compares the two.
#pragma once
#include "sampler_kernels.cuh"

namespace gphocs {

constexpr int kTeam = 8;                        // threads per locus in the team phase
constexpr int kTeamSlots = kThreads / kTeam;    // = kMaxBatchLoci
static_assert(kTeamSlots == kMaxBatchLoci, "one team per locus of a CTA batch");
constexpr int kSweepMaxNodes = 64;              // genealogies of up to 32 leaves (larger ones take the stepwise route)
// A proposal of these sweeps dirties one path to the root (coalescence time) or two paths that join (SPR): at most one
// result waits for its sibling's subtree, so ONE parking row per column is enough (k_eval keeps kStack for full
// evaluations); anything deeper would be re-read from the record just written, as in k_eval.
constexpr int kSweepStack = 1;
constexpr uint32_t kSweepHi = kSweepStack * kRow;
#ifndef GPHOCS_SWEEP_MINBLOCKS
#define GPHOCS_SWEEP_MINBLOCKS 8
#endif

// what the sweeps read of the model (SmpModel carries 39 populations and 32 bands: 2.7 KB per CTA)
struct SweepModel {
  int Q;
  int father[kSmpMaxPops], leavesBelow[kSmpMaxPops];
  unsigned long long below[kSmpMaxPops];
  double theta[kSmpMaxPops], tau[kSmpMaxPops], coalRate[kSmpMaxPops];
};

struct SweepSmem {
  // per locus slot: what lives for the whole sweep ...
  uint32_t offAge, offNode, offNeed, offPop, offCoal, offNcoal, perLocus;
  // ... and the scheduling scratch of one step, which shares its space with the column stack (never live together)
  uint32_t offSize, offWalk, perScratch;
  uint32_t perSched;
  // CTA regions
  uint32_t offLoci, offSched, offStack, offWords, offList, offTerm, offMeta, offProp, offCells, offModel, offLca, total;
  int W32;
};
__host__ __device__ inline SweepSmem sweepSmemLayout(int n, int maxLoci, int Q) {
  const int N = 2 * n - 1, NI = n - 1;
  SweepSmem m;
  m.W32 = (n + 7) / 8;
  m.offAge = 0;                                             // [N] double
  m.offCoal = m.offAge + (uint32_t)N * 8;                   // [Q] double
  m.offNode = m.offCoal + (uint32_t)Q * 8;                  // [N] NodeRec
  m.offNcoal = m.offNode + (uint32_t)N * sizeof(NodeRec);   // [Q] int
  m.offNeed = m.offNcoal + (uint32_t)Q * 4;                 // [N] uint8
  m.offPop = m.offNeed + (uint32_t)N;                       // [N] uint8
  m.perLocus = (m.offPop + (uint32_t)N + 15) & ~15u;
  m.offSize = 0;                                            // [NI] int
  m.offWalk = m.offSize + (uint32_t)NI * 4;                 // [N] uint32; the team phase keeps its lists here
  m.perScratch = (m.offWalk + (uint32_t)N * 4 + 15) & ~15u;
  m.perSched = (uint32_t)NI * sizeof(SchedEntryCompact);
  m.offLoci = 0;
  m.offSched = m.offLoci + m.perLocus * (uint32_t)maxLoci;
  // column stack (one parking row: lo halves, hi halves) = the root vectors after the walk = the scheduling scratch
  const uint32_t stackBytes = 2 * (uint32_t)kSweepHi, scratchBytes = m.perScratch * (uint32_t)maxLoci;
  m.offStack = m.offSched + m.perSched * (uint32_t)maxLoci;
  m.offWords = m.offStack + (stackBytes > scratchBytes ? stackBytes : scratchBytes);   // [W32][kThreads] uint32, whole sweep
  m.offList = m.offWords + (uint32_t)m.W32 * kThreads * 4;  // [maxLoci*NI] uint32: marked nodes of the batch
  m.offTerm = m.offList;                                    // [kThreads] double once the list is dead
  const uint32_t listBytes = (uint32_t)maxLoci * NI * 4, termBytes = (uint32_t)kThreads * 8;
  m.offMeta = (m.offList + (listBytes > termBytes ? listBytes : termBytes) + 15) & ~15u;   // per-slot scalars
  m.offProp = m.offMeta + (uint32_t)kTeamSlots * 64;        // [slots] SmpProposal
  m.offCells = (m.offProp + (uint32_t)kTeamSlots * sizeof(SmpProposal) + 15) & ~15u;   // list count, acceptance counters
  m.offModel = m.offCells + 16;
  m.offLca = m.offModel + (uint32_t)sizeof(SweepModel);     // [Q][Q] uint8: lowest population above (or equal to) both
  m.total = (m.offLca + (uint32_t)(Q * Q) + 15) & ~15u;
  return m;
}
__host__ __device__ inline size_t sweepSmemBytes(int n, int maxLoci, int Q) { return sweepSmemLayout(n, maxLoci, Q).total; }

struct Team {
  int j;           // 0..7 inside the team; 0 = leader
  int leader;      // lane of the leader inside the warp
  unsigned mask;   // the team's lanes
};
__device__ __forceinline__ double teamBcast(const Team& tm, double v) { return __shfl_sync(tm.mask, v, tm.leader); }
__device__ __forceinline__ int teamBcast(const Team& tm, int v) { return __shfl_sync(tm.mask, v, tm.leader); }
__device__ __forceinline__ int teamSumI(const Team& tm, int v) {
#pragma unroll
  for (int off = kTeam / 2; off > 0; off >>= 1) v += __shfl_xor_sync(tm.mask, v, off);
  return v;
}

// What the 32 lanes of the stepwise route's warp hold is held here by 8 threads: thread j stands for the "virtual
// lanes" j, j+8, j+16, j+24.  warpSumD's shuffle tree adds lanes that differ in bit 4, then bit 3, ..., bit 0; the first
// two levels are thread-local here, the last three cross the team — same additions, same order, same bits.
__device__ __forceinline__ double teamSumLikeWarp(const Team& tm, const double (&vl)[4]) {
  const double a = vl[0] + vl[2], b = vl[1] + vl[3];
  double v = a + b;
#pragma unroll
  for (int off = kTeam / 2; off > 0; off >>= 1) v += __shfl_xor_sync(tm.mask, v, off);
  return v;
}

// members of a team append the items they hold (`mine` = one bit per item, item k of thread j = id(j, k)) to a list in
// shared memory, densely; returns the list length.  Two rounds of ballots cover up to 8 items per thread.
template <typename IdOf>
__device__ __forceinline__ int teamCompact(const Team& tm, unsigned mine, int itemsPerThread, uint8_t* list, IdOf id) {
  int count = 0;
  for (int k = 0; k < itemsPerThread; k++) {
    const bool in = (mine >> k) & 1u;
    const unsigned ballot = (__ballot_sync(tm.mask, in) >> tm.leader) & ((1u << kTeam) - 1u);
    if (in) list[count + __popc(ballot & ((1u << tm.j) - 1u))] = (uint8_t)id(tm.j, k);
    count += __popc(ballot);
  }
  __syncwarp(tm.mask);
  return count;
}

// wlPopStat (sampler_kernels.cuh) on a team: coal statistic of population `pop` from the coalescences assigned to it.
// scratch: N bytes of the team's own shared memory (the ids of the population's coalescences).
__device__ inline double teamPopStat(const Team& tm, const SweepModel& m, const double* age, const uint8_t* np, int n, int N, int pop,
                                     int nStart, uint8_t* scratch) {
  unsigned mine = 0;
  const int per = (N - n + kTeam - 1) / kTeam;
  for (int k = 0; k < per; k++) {
    const int x = n + tm.j + kTeam * k;
    if (x < N && np[x] == pop) mine |= 1u << k;
  }
  const int total = teamCompact(tm, mine, per, scratch, [&](int j, int k) { return n + j + kTeam * k; });
  const double tau = m.tau[pop];
  const double end = m.father[pop] >= 0 ? m.tau[m.father[pop]] : kOldAge;
  if (total == 0) return (double)(nStart * (nStart - 1)) * (end - tau);
  const int R = N <= 32 ? 1 : 2;   // nodes per lane of the stepwise route
  double vl[4];
#pragma unroll
  for (int kk = 0; kk < 4; kk++) {
    double v = 0.0;
#pragma unroll 1
    for (int r = 0; r < R; r++) {
      const int x = tm.j + kTeam * kk + 32 * r;
      if (x < n || x >= N || np[x] != pop) continue;
      const double ax = age[x];
      int cnt = 0;
      double prev = tau;
#pragma unroll 1
      for (int i = 0; i < total; i++) {   // order-free: a count and a maximum
        const int y = scratch[i];
        const double ay = age[y];
        if (ay < ax || (ay == ax && y < x)) { cnt++; prev = fmax(prev, ay); }
      }
      const int lin = nStart - cnt;
      v += (double)(lin * (lin - 1)) * (ax - prev);
      if (cnt == total - 1) {
        const int restLin = lin - 1;
        v += (double)(restLin * (restLin - 1)) * (end - ax);
      }
    }
    vl[kk] = v;
  }
  return teamSumLikeWarp(tm, vl);
}

// smpAgeProposeBody on a team; the genealogy is the shared-memory copy behind `t`, np / coal / ncoal the slot's arrays
__device__ inline SmpProposal teamAgePropose(const Team& tm, const SweepModel& m, const TreeView& t, const uint8_t* np, const double* coal,
                                             const int* ncoal, int l, int n, int N, int inode, double finetune, unsigned long long seed,
                                             unsigned long long step, uint8_t* scratch) {
  SmpProposal pr = smpNoProposal();
  pr.node = inode;
  const int root = *t.root;
  if (root < n) return pr;
  double tnew = 0.0;
  int valid = 0, pop = 0, nStart = 0;
  if (tm.j == 0) {
    pop = np[inode];
    const double told = t.age[inode];
    const NodeRec rec = t.node[inode];
    const double lo = fmax(m.tau[pop], fmax(t.age[rec.left], t.age[rec.right]));
    double hi = m.father[pop] >= 0 ? m.tau[m.father[pop]] : kOldAge;
    if (inode != root) hi = fmin(hi, t.age[rec.father]);
    SmpRng rng(seed, (unsigned long long)l, step);
    tnew = smpReflect(told + finetune * rng.normal2(), lo, hi);
    valid = fabs(tnew - told) >= 1e-15;   // GPhoCS.c:2354-2358
    if (valid) {
      adjustAge(t, inode, tnew);
      nStart = m.leavesBelow[pop];
      for (int q = 0; q < m.Q; q++)
        if (q != pop && ((m.below[pop] >> q) & 1ull)) nStart -= ncoal[q];
    }
  }
  valid = teamBcast(tm, valid);
  if (!valid) return pr;
  pop = teamBcast(tm, pop);
  nStart = teamBcast(tm, nStart);
  __syncwarp(tm.mask);   // the leader's new age is in shared memory
  const double coalNew = teamPopStat(tm, m, t.age, np, n, N, pop, nStart, scratch);
  pr.genDelta = -(coalNew - coal[pop]) / m.theta[pop];
  pr.aux = coalNew;
  pr.pop = pop;
  pr.valid = 1;
  return pr;
}

// smpSprProposeBody on a team: every thread rings the clocks of its share of the branches
// (*oldGrandpa = the pruned father's father before the move, -1 if none: the second path of nodes to recompute starts there)
__device__ inline SmpProposal teamSprPropose(const Team& tm, const SweepModel& m, const TreeView& t, uint8_t* np, int l, int n, int N,
                                             int node, unsigned long long seed, unsigned long long step, int* oldGrandpa,
                                             const uint8_t* lca) {
  SmpProposal pr = smpNoProposal();
  *oldGrandpa = -1;
  const int root = *t.root;
  if (root < n || node == root) return pr;
  const int F = t.node[node].father;
  const NodeRec recF = t.node[F];
  const int S = recF.left + recF.right - node;
  const int G = recF.father;
  *oldGrandpa = G;
  const double t0 = t.age[node];
  const int pop0 = np[node];
  const double ageG = G >= 0 ? t.age[G] : kSmpInf;
  // (compacting the branches that can meet the pruned lineage before ringing their clocks was measured twice — with the
  // population walk repeated and with this table — and lost both times to the ballots it needs: 7.3 vs 7.0 ms)
  double bestT = kSmpInf;
  int bestX = -1, bestPop = -1;
  const SmpRng rng(seed, (unsigned long long)l, step);
#pragma unroll 1
  for (int x = tm.j; x < N; x += kTeam) {
    if (x == node || x == F) continue;
    const int fx = t.node[x].father;
    const double endx = x == S ? ageG : (fx >= 0 ? t.age[fx] : kSmpInf);
    int q = lca[np[x] * m.Q + pop0];   // first population in which the two lineages can meet: their common ancestor
    double sNow = fmax(fmax(t0, t.age[x]), m.tau[q]);
    if (sNow >= endx) continue;
    while (m.father[q] >= 0 && m.tau[m.father[q]] <= sNow) q = m.father[q];
    double need = rng.exponentialAt((unsigned long long)x);
#pragma unroll 1
    for (int it = 0; it < kSmpMaxPops; it++) {
      const double popEnd = m.father[q] >= 0 ? m.tau[m.father[q]] : kSmpInf;
      const double segEnd = fmin(endx, popEnd);
      const double rate = m.coalRate[q];
      if (rate * (segEnd - sNow) >= need) {
        const double T = sNow + need / rate;
        if (T < bestT) { bestT = T; bestX = x; bestPop = q; }
        break;
      }
      need -= rate * (segEnd - sNow);
      sNow = segEnd;
      if (sNow >= endx) break;
      q = m.father[q];
    }
  }
#pragma unroll
  for (int off = kTeam / 2; off > 0; off >>= 1) {   // earliest ring over the team (ties: lower node id)
    const double oT = __shfl_xor_sync(tm.mask, bestT, off);
    const int oX = __shfl_xor_sync(tm.mask, bestX, off);
    const int oP = __shfl_xor_sync(tm.mask, bestPop, off);
    if (oT < bestT || (oT == bestT && oX >= 0 && (bestX < 0 || oX < bestX))) { bestT = oT; bestX = oX; bestPop = oP; }
  }
  if (bestX >= 0) {
    __syncwarp(tm.mask);   // every thread of the team has read the genealogy it is about to see rewired
    if (tm.j == 0) {
      pr.pop = np[F];
      pr.node = F;
      spr(t, node, bestX, bestT);
      np[F] = (uint8_t)bestPop;
    }
    pr.valid = 1;   // the statistics are refreshed once, after the sweep (teamStatsAll at the end of k_sweep)
  }
  return pr;
}

// smpResolve on a team; returns 1 if the proposal counts as accepted
__device__ inline int teamResolve(const Team& tm, const TreeView& t, uint8_t* np, double* coal, const SmpProposal& pr, int l, int N,
                                  int kind, unsigned long long seed, unsigned long long step) {
  int ok = 0;
  if (pr.valid) {
    if (tm.j == 0) {
      const double lnacc = (*t.lnL - *t.savedLnL) + pr.genDelta;
      ok = lnacc >= 0.0;
      if (!ok) {
        SmpRng rng(seed, (unsigned long long)l, step);
        ok = rng.uniform() < exp(lnacc);
      }
    }
    ok = teamBcast(tm, ok);
    if (ok) {
      for (int x = tm.j; x < N; x += kTeam)
        if (t.node[x].flags & (F_RECALC | F_SAVED)) commitNode(t, x);   // both are no-ops on unmarked nodes
      if (tm.j == 0) {
        commitLocus(t);
        if (kind == 0) coal[pr.pop] = pr.aux;
      }
    } else {
      for (int x = tm.j; x < N; x += kTeam)
        if (t.node[x].flags & (F_RECALC | F_SAVED)) revertNode(t, x);
      if (tm.j == 0) {
        revertLocus(t);
        if (kind == 1) np[pr.node] = (uint8_t)pr.pop;
      }
    }
  } else if (kind == 0) {
    ok = 1;   // an unchanged age counts as accepted (GPhoCS.c:2354-2358)
  }
  __syncwarp(tm.mask);   // the proposal that follows reads what other threads of the team have just committed or reverted
  return ok;
}

// ------------------------------------------------------------------------------------------ what the sweep kernels share
// Shared-memory views, the column state of this thread and the phases every step runs after its team phase.
struct SweepCtx {
  unsigned char* smem;
  const SweepSmem* layp;   // the kernel's __grid_constant__ parameter: its fields are constant-bank operands
  Batch b;
  int tid, lane, warp, n, N, NI, nl, Q;
  // this thread's pattern column (of the first chunk, when the batch is one locus wider than the CTA)
  bool oversized;
  bool live;
  int colSlot, ph, cnt;
  char* clvCol;
  uint32_t myStack, myWords;

  __device__ __forceinline__ unsigned char* locus(int s) const { return smem + layp->offLoci + layp->perLocus * s; }
  __device__ __forceinline__ unsigned char* scratch(int s) const { return smem + layp->offStack + layp->perScratch * s; }
  __device__ __forceinline__ double* age(int s) const { return reinterpret_cast<double*>(locus(s) + layp->offAge); }
  __device__ __forceinline__ double* coal(int s) const { return reinterpret_cast<double*>(locus(s) + layp->offCoal); }
  __device__ __forceinline__ NodeRec* node(int s) const { return reinterpret_cast<NodeRec*>(locus(s) + layp->offNode); }
  __device__ __forceinline__ int* ncoal(int s) const { return reinterpret_cast<int*>(locus(s) + layp->offNcoal); }
  __device__ __forceinline__ uint8_t* need(int s) const { return locus(s) + layp->offNeed; }
  __device__ __forceinline__ uint8_t* pop(int s) const { return locus(s) + layp->offPop; }
  __device__ __forceinline__ int* size(int s) const { return reinterpret_cast<int*>(scratch(s) + layp->offSize); }
  __device__ __forceinline__ uint32_t* walk(int s) const { return reinterpret_cast<uint32_t*>(scratch(s) + layp->offWalk); }
  __device__ __forceinline__ SchedEntryCompact* sched(int s) const {
    return reinterpret_cast<SchedEntryCompact*>(smem + layp->offSched + layp->perSched * s);
  }
  __device__ __forceinline__ double* root4() const { return reinterpret_cast<double*>(smem + layp->offStack); }   // after the walk
  __device__ __forceinline__ double* term() const { return reinterpret_cast<double*>(smem + layp->offTerm); }
  // per-slot scalars
  __device__ __forceinline__ double* mRate() const { return reinterpret_cast<double*>(smem + layp->offMeta); }
  __device__ __forceinline__ double* mLnL() const { return mRate() + kTeamSlots; }
  __device__ __forceinline__ double* mSavedLnL() const { return mLnL() + kTeamSlots; }
  __device__ __forceinline__ unsigned long long* mEvals() const { return reinterpret_cast<unsigned long long*>(mSavedLnL() + kTeamSlots); }
  __device__ __forceinline__ unsigned long long* mEvalBytes() const { return mEvals() + kTeamSlots; }
  __device__ __forceinline__ int* mColStart() const { return reinterpret_cast<int*>(mEvalBytes() + kTeamSlots); }
  __device__ __forceinline__ int* mP() const { return mColStart() + kTeamSlots; }
  __device__ __forceinline__ int* mK() const { return mP() + kTeamSlots; }
  __device__ __forceinline__ int* mRoot() const { return mK() + kTeamSlots; }
  __device__ __forceinline__ int* mSavedRoot() const { return mRoot() + kTeamSlots; }
  __device__ __forceinline__ int* mActive() const { return mSavedRoot() + kTeamSlots; }
  __device__ __forceinline__ SmpProposal* prop() const { return reinterpret_cast<SmpProposal*>(smem + layp->offProp); }
  // not inside the list region: the root terms take that over while the count is being reset
  __device__ __forceinline__ int* listCount() const { return reinterpret_cast<int*>(smem + layp->offCells); }
  __device__ __forceinline__ unsigned int* accepted() const { return reinterpret_cast<unsigned int*>(smem + layp->offCells) + 1; }   // [2]
  __device__ __forceinline__ uint32_t* list() const { return reinterpret_cast<uint32_t*>(smem + layp->offList); }
  __device__ __forceinline__ SweepModel& model() const { return *reinterpret_cast<SweepModel*>(smem + layp->offModel); }
  __device__ __forceinline__ uint8_t* lca() const { return smem + layp->offLca; }
};

// stage the batch: model, per-locus scalars, genealogies, population assignments, coal statistics, leaf codes; ends
// with a barrier
__device__ inline void sweepStage(SweepCtx& c, unsigned char* smem, const SweepSmem& lay, const StoreDev& d, const SmpDev& sd,
                                  const SmpModel* __restrict__ mp, const Batch& b) {
  c.smem = smem; c.layp = &lay; c.b = b;
  c.tid = threadIdx.x; c.lane = c.tid & 31; c.warp = c.tid >> 5;
  c.n = d.n; c.N = d.N; c.NI = d.NI; c.nl = b.numLoci; c.Q = sd.Q;
  const int tid = c.tid, Q = c.Q, N = c.N, nl = c.nl;
  SweepModel& sModel = c.model();
  if (tid == 0) sModel.Q = Q;
  for (int p = tid; p < Q; p += kThreads) {
    sModel.father[p] = mp->father[p]; sModel.leavesBelow[p] = mp->leavesBelow[p]; sModel.below[p] = mp->below[p];
    sModel.theta[p] = mp->theta[p]; sModel.tau[p] = mp->tau[p]; sModel.coalRate[p] = mp->coalRate[p];
  }
  for (int i = tid; i < Q * Q; i += kThreads) {   // lowest common population of every pair (the SPR proposal asks per branch)
    const int pa = i / Q, pb = i - pa * Q;
    int q = pa;
    while (!((mp->below[q] >> pb) & 1ull)) q = mp->father[q];
    c.lca()[i] = (uint8_t)q;
  }
  unsigned long long w0 = 0ull, w1 = 0ull;
  c.ph = 0; c.cnt = 0;
  c.oversized = b.scratchOff >= 0;
  c.live = tid < b.numCols;
  if (c.live) {
    const int col = b.firstCol + tid;
    w0 = d.leafWords[col];
    if (d.W > 1) w1 = d.leafWords[(size_t)d.Ct + col];
    if (!c.oversized) { c.ph = d.grpPhases[col]; c.cnt = d.grpCount[col]; }
  }
  if (tid < nl) {
    const int l = b.firstLocus + tid;
    const int c0 = d.colStart[l];
    c.mColStart()[tid] = c0;
    c.mP()[tid] = d.colStart[l + 1] - c0;
    const int root = d.root[l];
    c.mRoot()[tid] = root;
    c.mSavedRoot()[tid] = d.savedRoot[l];
    c.mRate()[tid] = d.rate[l];
    c.mActive()[tid] = (c.mP()[tid] > 0) && (root >= c.n);
    c.mK()[tid] = 0;
    c.mLnL()[tid] = d.lnL[l];
    c.mSavedLnL()[tid] = d.savedLnL[l];
    c.mEvals()[tid] = 0ull;
    c.mEvalBytes()[tid] = 0ull;
  }
  if (tid < 2) c.accepted()[tid] = 0u;
  if (tid == 0) *c.listCount() = 0;
  for (int s = c.warp; s < nl; s += kWarps) {
    const int l = b.firstLocus + s;
    const size_t g0 = (size_t)l * N;
    NodeRec* nd = c.node(s);
    double* age = c.age(s);
    uint8_t* need = c.need(s);
    uint8_t* pop = c.pop(s);
    for (int v = c.lane; v < N; v += 32) {
      nd[v] = d.node[g0 + v];
      age[v] = d.age[g0 + v];
      pop[v] = sd.nodePop[g0 + v];
      need[v] = 0;
    }
    for (int p = c.lane; p < Q; p += 32) {
      c.coal(s)[p] = sd.coal[(size_t)l * Q + p];
      c.ncoal(s)[p] = sd.ncoal[(size_t)l * Q + p];
    }
  }
  c.myStack = smemAddr(smem + lay.offStack) + tid * 16;
  c.myWords = smemAddr(smem + lay.offWords) + tid * 4;
  if (c.live) {   // this column's leaf masks, 8 leaves per 32-bit word
    stsU32(c.myWords, (uint32_t)w0);
    if (lay.W32 > 1) stsU32(c.myWords + kThreads * 4, (uint32_t)(w0 >> 32));
    if (lay.W32 > 2) stsU32(c.myWords + 2 * kThreads * 4, (uint32_t)w1);
    if (lay.W32 > 3) stsU32(c.myWords + 3 * kThreads * 4, (uint32_t)(w1 >> 32));
  }
  __syncthreads();
  c.colSlot = 0;   // the locus this thread's column belongs to
  if (c.live) {
    const int col = b.firstCol + tid;
    while (c.colSlot + 1 < nl && col >= c.mColStart()[c.colSlot + 1]) c.colSlot++;
  }
  c.clvCol = reinterpret_cast<char*>(d.clv + (size_t)c.mColStart()[c.colSlot] * c.NI * 8 +
                                     (size_t)(c.live ? b.firstCol + tid - c.mColStart()[c.colSlot] : 0) * 4);
}

// the genealogy of a slot as the edit protocol sees it: current arrays in shared memory, saved copies in HBM
__device__ __forceinline__ TreeView sweepTreeView(const SweepCtx& c, const StoreDev& d, int slot) {
  TreeView t;
  const size_t o = (size_t)(c.b.firstLocus + slot) * c.N;
  t.node = c.node(slot); t.saved = d.saved + o;
  t.age = c.age(slot); t.svAge = d.svAge + o;
  t.root = c.mRoot() + slot; t.savedRoot = c.mSavedRoot() + slot;
  t.lnL = c.mLnL() + slot; t.savedLnL = c.mSavedLnL() + slot; t.rate = c.mRate() + slot;
  t.numLeaves = c.n;
  t.numPatterns = c.mP()[slot];
  return t;
}

// end of a team phase that has made a proposal for `slot`: k_eval phases B (dirty nodes and their ancestors; a moved
// leaf dirties its father, .c:1569-1575) and C0 (marked nodes of the batch compacted into one list, their destination
// buffers flipped)
__device__ inline void sweepMarkAndCompact(const SweepCtx& c, const Team& tm, int slot) {
  const int n = c.n, N = c.N;
  if (tm.j == 0) {
    c.mK()[slot] = 0;
    if (c.mActive()[slot]) c.mSavedLnL()[slot] = c.mLnL()[slot];   // what every evaluation starts with (.c:440)
  }
  __syncwarp(tm.mask);
  if (!c.mActive()[slot]) return;
  NodeRec* nd = c.node(slot);
  uint8_t* need = c.need(slot);
  for (int v = tm.j; v < N; v += kTeam)
    if (nd[v].flags & F_RECALC) {
      int u = v < n ? nd[v].father : v;
      for (int k = 0; u >= 0 && !need[u] && k < N; k++) {   // concurrent walkers store the same 1 (see k_eval)
        need[u] = 1;
        u = nd[u].father;
      }
    }
  __syncwarp(tm.mask);
  for (int v0 = n; v0 < N; v0 += kTeam) {
    const int v = v0 + tm.j;
    const bool marked = v < N && need[v];
    const unsigned ballot = (__ballot_sync(tm.mask, marked) >> tm.leader) & ((1u << kTeam) - 1u);
    int base = 0;
    if (tm.j == 0 && ballot) base = atomicAdd(c.listCount(), __popc(ballot));
    base = teamBcast(tm, base);
    if (marked) {
      c.size(slot)[v - n] = 0;   // phase C counts into it (its space was the column stack a moment ago)
      c.list()[base + __popc(ballot & ((1u << tm.j) - 1u))] = (uint32_t)(slot << 16 | v);
      const uint8_t f = nd[v].flags;
      if (!(f & F_RECALC)) nd[v].flags = (uint8_t)((f ^ F_SEL) | F_RECALC);
    }
  }
}

// Column and root phases for a batch that is ONE locus with more pattern columns than the CTA has threads: the columns
// are walked in chunks of kThreads, root vectors go through rootScratch, the terms are reduced by the fixed-order block
// tree of k_eval's oversized path (same bits).  Ends with a barrier.
__device__ inline void sweepEvaluateWide(const SweepCtx& c, const StoreDev& d, int listCount) {
  const int tid = c.tid;
  const uint32_t* sList = c.list();
  const int k = c.mActive()[0] ? c.mK()[0] : 0;
  const int P = c.mP()[0], c0 = c.mColStart()[0];
  if (k > 0) {
    const int numChunks = (c.b.numCols + kThreads - 1) / kThreads;
    for (int chunk = 0; chunk < numChunks; chunk++) {
      const int p = chunk * kThreads + tid;
      if (p >= c.b.numCols) break;
      const int col = c.b.firstCol + p;
      const unsigned long long w0 = d.leafWords[col];
      const unsigned long long w1 = d.W > 1 ? d.leafWords[(size_t)d.Ct + col] : 0ull;
      stsU32(c.myWords, (uint32_t)w0);
      if (c.layp->W32 > 1) stsU32(c.myWords + kThreads * 4, (uint32_t)(w0 >> 32));
      if (c.layp->W32 > 2) stsU32(c.myWords + 2 * kThreads * 4, (uint32_t)w1);
      if (c.layp->W32 > 3) stsU32(c.myWords + 3 * kThreads * 4, (uint32_t)(w1 >> 32));
      char* clvCol = reinterpret_cast<char*>(d.clv + (size_t)c0 * c.NI * 8 + (size_t)p * 4);
      double pv[4] = {0.0, 0.0, 0.0, 0.0};
      columnWalk<kSweepHi, true>(smemAddr(c.sched(0)), k, clvCol, c.myStack, c.myWords, true, pv);
      double* dst = d.rootScratch + ((size_t)c.b.scratchOff + p) * 4;
#pragma unroll
      for (int q = 0; q < 4; q++) dst[q] = pv[q];
    }
  }
  __syncthreads();
  for (int j = tid; j < listCount; j += kThreads) c.need(sList[j] >> 16)[sList[j] & 0xffff] = 0;
  __syncthreads();   // the list is dead: its space takes the terms
  double acc = 0.0;
  if (k > 0) {
    const double* src = d.rootScratch + (size_t)c.b.scratchOff * 4;
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
  double* sTerm = c.term();
  sTerm[tid] = acc;
  if (tid == 0) *c.listCount() = 0;
  __syncthreads();
  for (int off = kThreads / 2; off > 0; off >>= 1) {
    if (tid < off) sTerm[tid] += sTerm[tid + off];
    __syncthreads();
  }
  if (tid == 0 && k > 0) {
    c.mLnL()[0] = sTerm[0];
    c.mEvals()[0]++;
    c.mEvalBytes()[0] += 32ull * (unsigned long long)P * (2ull * (unsigned long long)k + 1ull);
  }
  __syncthreads();
}

// schedules for arbitrary sets of marked nodes (k_global_move): k_eval phases C, D over the list sweepMarkAndCompact left;
// returns the list length
__device__ inline int sweepListSchedule(const SweepCtx& c) {
  const int tid = c.tid, n = c.n, N = c.N;
  __syncthreads();
  const int listCount = *c.listCount();
  const uint32_t* sList = c.list();
  // ---- k_eval phase C: marked nodes per subtree
  for (int j = tid; j < listCount; j += kThreads) {
    const int s = sList[j] >> 16, v = sList[j] & 0xffff;
    const NodeRec* nd = c.node(s);
    int* size = c.size(s);
    int a = v;
    for (int k = 0; a >= 0 && k < N; k++) {
      atomicAdd(&size[a - n], 1);
      a = nd[a].father;
    }
  }
  __syncthreads();
  // ---- k_eval phase D1: what a node adds to the post-order start of everything below it (heavier child first)
  for (int j = tid; j < listCount; j += kThreads) {
    const int s = sList[j] >> 16, v = sList[j] & 0xffff;
    const NodeRec* nd = c.node(s);
    const uint8_t* need = c.need(s);
    const int* size = c.size(s);
    const int a = nd[v].father;
    uint32_t contrib = 0;
    if (a >= 0) {
      const int l = nd[a].left, r = nd[a].right;
      const int wl = (l >= n && need[l]) ? size[l - n] : 0;
      const int wr = (r >= n && need[r]) ? size[r - n] : 0;
      const int first = wl >= wr ? l : r;
      if (v != first) contrib = (uint32_t)(v == l ? wr : wl);
    }
    c.walk(s)[v] = (uint32_t)(a + 1) | (contrib << 16);
  }
  __syncthreads();
  // ---- k_eval phase D2: position, stack depth, child sources and JC69 edge terms of every marked node
  for (int j = tid; j < listCount; j += kThreads) {
    const int s = sList[j] >> 16, v = sList[j] & 0xffff;
    const NodeRec* nd = c.node(s);
    const uint8_t* need = c.need(s);
    const int* size = c.size(s);
    const uint32_t* walk = c.walk(s);
    const double* age = c.age(s);
    const double rate = c.mRate()[s];
    const int rootId = c.mRoot()[s];
    int start = 0, depth = 0;
    {
      uint32_t w = walk[v];
      for (int k = 0; k < N; k++) {
        const uint32_t cc = w >> 16;
        start += cc;
        depth += cc != 0;
        const int a = (int)(w & 0xffffu) - 1;
        if (a < 0) break;
        w = walk[a];
      }
    }
    const int l = nd[v].left, r = nd[v].right;
    const int wl = (l >= n && need[l]) ? size[l - n] : 0;
    const int wr = (r >= n && need[r]) ? size[r - n] : 0;
    const bool leftFirst = wl >= wr;
    const int A = leftFirst ? l : r, B = leftFirst ? r : l;
    const int wA = leftFirst ? wl : wr, wB = leftFirst ? wr : wl;
    const uint32_t strideBytes = (uint32_t)c.mP()[s] * 32u;
    auto record = [&](int x) { return (uint32_t)((x - n) * 2 + (nd[x].flags & F_SEL)) * strideBytes; };
    auto leafRef = [&](int x) { return (uint32_t)(x >> 3) * (kThreads * 4u) | ((uint32_t)(x & 7) * 4u) << 16; };
    uint32_t kindA, kindB, offA = 0, offB = 0;
    if (A < n) { kindA = SRC_LEAF; offA = leafRef(A); }
    else if (wA > 0 && wB == 0) { kindA = SRC_TOP; }
    else if (wA > 0 && depth < kSweepStack) { kindA = SRC_STACK; offA = ((uint32_t)depth * kRow) << 16; }
    else { kindA = SRC_GLOBAL; offA = record(A); }
    if (B < n) { kindB = SRC_LEAF; offB = leafRef(B); }
    else if (wB > 0) { kindB = SRC_TOP; }
    else { kindB = SRC_GLOBAL; offB = record(B); }
    SchedEntryCompact en;
    const double av = age[v];
    en.e0A = edgeProb(rate * (av - age[A]));
    en.e0B = edgeProb(rate * (av - age[B]));
    en.offA = offA; en.offB = offB;
    en.dstOff = record(v);
    uint32_t push = 0xffffu;
    {
      const int f = nd[v].father;
      if (f >= 0 && v != rootId && depth < kSweepStack) {
        const int fl = nd[f].left, fr = nd[f].right;
        const int wfl = (fl >= n && need[fl]) ? size[fl - n] : 0;
        const int wfr = (fr >= n && need[fr]) ? size[fr - n] : 0;
        const int first = wfl >= wfr ? fl : fr;
        const int wSibling = v == fl ? wfr : wfl;
        if (v == first && wSibling > 0) push = (uint32_t)depth * kRow;
      }
    }
    en.ctl = kindA | (kindB << 2) | (push << 16);
    c.sched(s)[start + size[v - n] - 1] = en;
    if (v == rootId) c.mK()[s] = size[v - n];
  }
  return listCount;
}

// the rest of a step, for the whole CTA, once every locus' schedule stands: k_eval phases E (column walk) and F (root);
// leaves the new log-likelihoods in mLnL and ends with a barrier.  listCount: entries of the marked-node list whose
// dirty marks are to be cleared (0 when the schedules were built without the list).
__device__ inline void sweepWalkAndRoot(const SweepCtx& c, const StoreDev& d, int listCount) {
  const int tid = c.tid;
  const uint32_t* sList = c.list();
  __syncthreads();
  if (c.oversized) { sweepEvaluateWide(c, d, listCount); return; }
  // ---- column phase (k_eval phase E)
  double pv[4] = {0.0, 0.0, 0.0, 0.0};
  const int k = (c.live && c.mActive()[c.colSlot]) ? c.mK()[c.colSlot] : 0;
  if (k > 0) columnWalk<kSweepHi, true>(smemAddr(c.sched(c.colSlot)), k, c.clvCol, c.myStack, c.myWords, true, pv);
  __syncthreads();   // the stack is dead; its space takes the root vectors
  double* sRoot = c.root4();
#pragma unroll
  for (int q = 0; q < 4; q++) sRoot[tid * 4 + q] = pv[q];
  for (int j = tid; j < listCount; j += kThreads) {   // dirty marks back to zero for the next step
    const int s = sList[j] >> 16, v = sList[j] & 0xffff;
    c.need(s)[v] = 0;
  }
  __syncthreads();
  // ---- root phase (k_eval phase F): 4*phases conditionals per phase group in the reference's order (.c:470-479)
  double term = 0.0;
  if (k > 0 && c.ph > 0) {
    double prob = 0.0;
    const int numConds = 4 * c.ph;
    for (int j = 0; j < numConds; j++) prob += sRoot[tid * 4 + j];
    term = log(prob / numConds) * c.cnt;
  }
  c.term()[tid] = term;
  if (tid == 0) *c.listCount() = 0;
  __syncthreads();
  if (tid < c.nl && c.mActive()[tid] && c.mK()[tid] > 0) {   // per-locus sum in pattern order
    const int P = c.mP()[tid];
    const double* tt = c.term() + (c.mColStart()[tid] - c.b.firstCol);
    double lnl = 0.0;
    for (int j = 0; j < P; j++) lnl += tt[j];
    c.mLnL()[tid] = lnl;
    c.mEvals()[tid]++;
    c.mEvalBytes()[tid] += 32ull * (unsigned long long)P * (2ull * (unsigned long long)c.mK()[tid] + 1ull);
  }
  __syncthreads();
}

// the final state goes back: genealogies (every proposal is resolved: flags hold buffer selectors only), population
// assignments, log-likelihoods, coal statistics, acceptance counters (kinds 0 and 1), evaluation accounting
__device__ inline void sweepWriteBack(const SweepCtx& c, const StoreDev& d, const SmpDev& sd, const unsigned int (&acceptedMine)[2],
                                      bool leaderOn, bool keepProposals = false) {
  const int tid = c.tid, N = c.N, Q = c.Q, nl = c.nl;
  if (leaderOn) {
    if (acceptedMine[0]) atomicAdd(&c.accepted()[0], acceptedMine[0]);
    if (acceptedMine[1]) atomicAdd(&c.accepted()[1], acceptedMine[1]);
  }
  __syncthreads();
  for (int s = c.warp; s < nl; s += kWarps) {
    const int l = c.b.firstLocus + s;
    const size_t g0 = (size_t)l * N;
    const NodeRec* nd = c.node(s);
    const double* age = c.age(s);
    const uint8_t* pop = c.pop(s);
    for (int v = c.lane; v < N; v += 32) {
      d.node[g0 + v] = nd[v];
      d.age[g0 + v] = age[v];
      sd.nodePop[g0 + v] = pop[v];
    }
    for (int p = c.lane; p < Q; p += 32) { sd.coal[(size_t)l * Q + p] = c.coal(s)[p]; sd.ncoal[(size_t)l * Q + p] = c.ncoal(s)[p]; }
  }
  if (tid < nl) {
    const int l = c.b.firstLocus + tid;
    d.root[l] = c.mRoot()[tid];
    d.savedRoot[l] = c.mSavedRoot()[tid];
    d.lnL[l] = c.mLnL()[tid];
    d.savedLnL[l] = c.mSavedLnL()[tid];
    sd.prop[l] = keepProposals ? c.prop()[tid] : smpNoProposal();   // (a global move stays pending until the next launch)
  }
  if (tid < 2 && c.accepted()[tid]) atomicAdd(sd.accepted + tid, (unsigned long long)c.accepted()[tid]);
  if (d.evalCounters && c.warp == 0) {
    unsigned long long evals = c.lane < nl ? c.mEvals()[c.lane] : 0ull, evalBytes = c.lane < nl ? c.mEvalBytes()[c.lane] : 0ull;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      evals += __shfl_xor_sync(0xffffffffu, evals, off);
      evalBytes += __shfl_xor_sync(0xffffffffu, evalBytes, off);
    }
    if (c.lane == 0 && evals) { atomicAdd(d.evalCounters, evals); atomicAdd(d.evalCounters + 1, evalBytes); }
  }
}

// Schedule of a node move, built by the locus' team without the batch-wide list phases.  What a coalescence-time move or
// an SPR dirties is a path to the root (from `first`) or two paths that join (the second from `second`, the pruned
// father's old father): tail A = path 1 below the junction, tail B = path 2 below it, stem = junction to root.  The
// longer tail goes first and parks its top on the column stack's one row; every other recomputed child is the entry
// right before its father's (register top).  Children first, same arithmetic per node as k_eval's schedule — the
// product of the two children's factors does not depend on which of them is visited first — so the conditional vectors
// are k_eval's bit for bit.  Also flips the destination buffers of the path nodes (k_eval phase C0).
__device__ inline void sweepPathSchedule(const SweepCtx& c, const Team& tm, int slot, int first, int second) {
  const int n = c.n;
  __syncwarp(tm.mask);   // the team is done with what it kept in its scratch during the proposal
  if (tm.j == 0) {
    c.mK()[slot] = 0;
    if (c.mActive()[slot]) c.mSavedLnL()[slot] = c.mLnL()[slot];   // what every evaluation starts with (.c:440)
  }
  NodeRec* nd = c.node(slot);
  uint8_t* need = c.need(slot);
  uint8_t* ord = reinterpret_cast<uint8_t*>(c.walk(slot));   // team scratch: the nodes in schedule order
  // the leader walks the two paths (tmpA: first .. root; tmpB: second .. below the first node on path 1) ...
  uint8_t* tmpA = ord;
  uint8_t* tmpB = ord + c.NI;
  int a = 0, lenA = 0, lenB = 0;
  if (tm.j == 0 && c.mActive()[slot] && first >= n) {
    for (int u = first; u >= 0 && a < c.NI; u = nd[u].father) { tmpA[a++] = (uint8_t)u; need[u] = 1; }
    int bLen = 0;
    for (int u = second; u >= n && !need[u] && bLen < c.NI; u = nd[u].father) tmpB[bLen++] = (uint8_t)u;
    if (bLen > 0) {
      const int top = nd[tmpB[bLen - 1]].father;   // the junction: on path 1
      for (lenA = 0; lenA < a && tmpA[lenA] != top; lenA++) {}
    }
    lenB = bLen;
    c.mK()[slot] = a + bLen;
  }
  a = teamBcast(tm, a);
  lenA = teamBcast(tm, lenA);
  lenB = teamBcast(tm, lenB);
  const int k = a + lenB;
  __syncwarp(tm.mask);
  if (k == 0) return;
  // ... the team does the rest.  Order: longer tail, other tail, then the rest of path 1.
  const bool aFirst = lenA >= lenB;
  const int lenFirst = aFirst ? lenA : lenB;
  const bool bothTails = lenA > 0 && lenB > 0;
  auto nodeAt = [&](int i) -> int {
    if (i < lenFirst) return aFirst ? tmpA[i] : tmpB[i];
    if (i < lenA + lenB) return aFirst ? tmpB[i - lenA] : tmpA[i - lenB];
    return tmpA[i - lenB];
  };
  for (int i = tm.j; i < k; i += kTeam) {   // marks of path 2, destination buffers of every node to recompute
    const int v = nodeAt(i);
    need[v] = 1;
    const uint8_t f = nd[v].flags;
    if (!(f & F_RECALC)) nd[v].flags = (uint8_t)((f ^ F_SEL) | F_RECALC);
  }
  __syncwarp(tm.mask);
  const double* age = c.age(slot);
  const double rate = c.mRate()[slot];
  const uint32_t strideBytes = (uint32_t)c.mP()[slot] * 32u;
  auto record = [&](int x) { return (uint32_t)((x - n) * 2 + (nd[x].flags & F_SEL)) * strideBytes; };
  auto leafRef = [&](int x) { return (uint32_t)(x >> 3) * (kThreads * 4u) | ((uint32_t)(x & 7) * 4u) << 16; };
  for (int i = tm.j; i < k; i += kTeam) {
    const int v = nodeAt(i);
    const int prevNode = i > 0 ? nodeAt(i - 1) : -1;
    uint32_t kind[2], off[2];
    int child[2] = {nd[v].left, nd[v].right};
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const int x = child[s];
      off[s] = 0;
      if (x < n) { kind[s] = SRC_LEAF; off[s] = leafRef(x); }
      else if (!need[x]) { kind[s] = SRC_GLOBAL; off[s] = record(x); }   // clean child: its current buffer
      else if (x == prevNode) { kind[s] = SRC_TOP; }
      else { kind[s] = SRC_STACK; off[s] = 0u; }                          // the first tail's top, parked in row 0
    }
    SchedEntryCompact en;
    const double av = age[v];
    en.e0A = edgeProb(rate * (av - age[child[0]]));
    en.e0B = edgeProb(rate * (av - age[child[1]]));
    en.offA = off[0]; en.offB = off[1];
    en.dstOff = record(v);
    const uint32_t push = (bothTails && i == lenFirst - 1) ? 0u : 0xffffu;
    en.ctl = kind[0] | (kind[1] << 2) | (push << 16);
    c.sched(slot)[i] = en;
  }
  __syncwarp(tm.mask);
  for (int i = tm.j; i < k; i += kTeam) need[nodeAt(i)] = 0;   // marks back to zero for the next step
}

__device__ __forceinline__ Team sweepTeam(int tid) {
  Team tm;
  tm.j = tid & (kTeam - 1);
  tm.leader = (tid & 31) & ~(kTeam - 1);
  tm.mask = ((1u << kTeam) - 1u) << tm.leader;
  return tm;
}

// wlStats (sampler_kernels.cuh) on a team, with population A's split time overridden: coalT[p], ncoalT[p] for every p.
// terms: NI doubles of team scratch.  Per-population sums are formed in the order of wlStats' 32-lane shuffle tree
// (teamSumLikeWarp), so the statistics are those of the stepwise route bit for bit.
__device__ inline void teamStatsAll(const Team& tm, const SweepModel& m, const double* age, const uint8_t* np, int n, int N, int ovPop,
                                    double ovTau, double* terms, double* coalT, int* ncoalT) {
  const int Q = m.Q;
  auto tauOf = [&](int p) { return p == ovPop ? ovTau : m.tau[p]; };
  auto endOf = [&](int p) { return m.father[p] >= 0 ? tauOf(m.father[p]) : kOldAge; };
  auto entering = [&](int p) {   // lineages entering p = samples below it minus coalescences strictly below it
    int lin = m.leavesBelow[p];
    for (int q = 0; q < Q; q++)
      if (q != p && ((m.below[p] >> q) & 1ull)) lin -= ncoalT[q];
    return lin;
  };
  for (int p = tm.j; p < Q; p += kTeam) {   // coalescences per population
    int k = 0;
    for (int x = n; x < N; x++) k += np[x] == p;
    ncoalT[p] = k;
  }
  __syncwarp(tm.mask);
#pragma unroll 1
  for (int x = n + tm.j; x < N; x += kTeam) {
    const int p = np[x];
    const double ax = age[x];
    int cnt = 0;
    double prev = tauOf(p);
#pragma unroll 1
    for (int y = n; y < N; y++) {
      if (np[y] != p) continue;
      const double ay = age[y];
      if (ay < ax || (ay == ax && y < x)) { cnt++; prev = fmax(prev, ay); }
    }
    const int lin = entering(p) - cnt;
    double term = (double)(lin * (lin - 1)) * (ax - prev);
    if (cnt == ncoalT[p] - 1) {   // last coalescence of its population: the interval up to the population's end
      const int rest = lin - 1;
      term += (double)(rest * (rest - 1)) * (endOf(p) - ax);
    }
    terms[x - n] = term;
  }
  __syncwarp(tm.mask);
  const int R = N <= 32 ? 1 : 2;   // nodes per lane of the stepwise route
#pragma unroll 1
  for (int p = 0; p < Q; p++) {
    double vl[4];
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      double v = 0.0;
      for (int r = 0; r < R; r++) {
        const int x = tm.j + kTeam * kk + 32 * r;
        v += (x >= n && x < N && np[x] == p) ? terms[x - n] : 0.0;
      }
      vl[kk] = v;
    }
    double v = teamSumLikeWarp(tm, vl);
    if (ncoalT[p] == 0) {
      const int lin = entering(p);
      v = (double)(lin * (lin - 1)) * (endOf(p) - tauOf(p));
    }
    if (tm.j == 0) coalT[p] = v;
  }
  __syncwarp(tm.mask);
}

// ------------------------------------------------------------------------------------------ models without migration bands
__global__ void __launch_bounds__(kThreads, GPHOCS_SWEEP_MINBLOCKS)
k_sweep(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, const Batch* __restrict__ batches, const __grid_constant__ SweepSmem lay,
        double ftCoal, unsigned long long seed, unsigned long long step0) {
  extern __shared__ __align__(16) unsigned char smem[];
  SweepCtx c;
  sweepStage(c, smem, lay, d, sd, mp, batches[blockIdx.x]);
  const SweepModel& m = c.model();
  const int n = c.n, N = c.N;
  const Team tm = sweepTeam(c.tid);
  const int slot = c.tid / kTeam;
  const bool teamOn = slot < c.nl;
  const int myLocus = c.b.firstLocus + slot;
  TreeView t;
  if (teamOn) t = sweepTreeView(c, d, slot);
  unsigned int accepted[2] = {0u, 0u};

  const int numAge = ftCoal > 0.0 ? c.NI : 0, numSteps = numAge + N;
  for (int it = 0; it <= numSteps; it++) {
    // ---- team phase: the previous proposal is resolved, the next one made
    if (teamOn) {
      if (it > 0) {
        const int kind = it - 1 < numAge ? 0 : 1;
        const int ok = teamResolve(tm, t, c.pop(slot), c.coal(slot), c.prop()[slot], myLocus, N, kind, seed, step0 + 2ull * (it - 1) + 1ull);
        if (tm.j == 0) accepted[kind] += ok;
      }
      if (it < numSteps) {
        const unsigned long long step = step0 + 2ull * it;
        uint8_t* scratch = reinterpret_cast<uint8_t*>(c.walk(slot));   // team scratch: N words
        int second = -1;
        const SmpProposal pr = it < numAge
            ? teamAgePropose(tm, m, t, c.pop(slot), c.coal(slot), c.ncoal(slot), myLocus, n, N, n + it, ftCoal, seed, step, scratch)
            : teamSprPropose(tm, m, t, c.pop(slot), myLocus, n, N, it - numAge, seed, step, &second, c.lca());
        if (tm.j == 0) c.prop()[slot] = pr;
        // nodes to recompute: the moved node (coalescence time) or the moved father and its old father (SPR), and
        // their ancestors; nothing if no proposal was made
        const int first = pr.valid ? (it < numAge ? n + it : teamBcast(tm, pr.node)) : -1;
        sweepPathSchedule(c, tm, slot, first, pr.valid ? second : -1);
      }
    }
    if (it == numSteps) break;
    sweepWalkAndRoot(c, d, 0);
  }
  // ---- statistics of the final genealogy (the SPR sweep moved coalescences between populations): k_smp_init_stats
  if (teamOn && *t.root >= n) {
    double* terms = reinterpret_cast<double*>(c.sched(slot));   // team scratch, as in k_global_move
    double* coalT = terms + c.NI;
    int* ncoalT = reinterpret_cast<int*>(c.walk(slot));
    teamStatsAll(tm, m, t.age, c.pop(slot), n, N, -1, 0.0, terms, coalT, ncoalT);
    for (int p = tm.j; p < c.Q; p += kTeam) { c.coal(slot)[p] = coalT[p]; c.ncoal(slot)[p] = ncoalT[p]; }
  }
  sweepWriteBack(c, d, sd, accepted, teamOn && tm.j == 0);
}

// ------------------------------------------------------------------------------------------ global moves, one launch each
// UpdateTau (GPhoCS.c:3224-3990, with the rubber band of patch.c:596-801), UpdateSampleAge (GPhoCS.c:4006) and mixing
// (GPhoCS.c:4688) propose one change for ALL loci and are accepted or rejected as a whole, from sums over the loci.  On
// the stepwise route that is five launches per move (resolve the previous move, propose, evaluate, reduce, reduce).
// Here one launch does it with the machinery of k_sweep: the CTA stages its batch, a team per locus first resolves
// the PREVIOUS global move (the host knows its outcome by now and passes it in), then makes this move's proposal —
// rubber band + the statistics under the proposed split time, or the rescaling of every age — the list / column /
// root phases evaluate it.  The proposal records and log-likelihoods go back per locus and k_smp_reduce sums them
// exactly as on the stepwise route: two launches per move, and the same chain bit for bit.  Models without migration
// bands.
struct GlobalMove {
  int prevKind;       // -1: nothing pending; 0: a split-time / sample-age move; 1: mixing
  int prevAccept;
  double prevC;       // mixing: the factor of the pending move
  int kind;           // -1: resolve only; 0: split-time / sample-age move of population A; 1: mixing by c
  int A;
  double tauOld, tauNew, lb, ub, f0, f1, c;
};

__global__ void __launch_bounds__(kThreads, GPHOCS_SWEEP_MINBLOCKS)
k_global_move(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, const Batch* __restrict__ batches, const __grid_constant__ SweepSmem lay,
              const __grid_constant__ GlobalMove gm) {
  extern __shared__ __align__(16) unsigned char smem[];
  SweepCtx c;
  sweepStage(c, smem, lay, d, sd, mp, batches[blockIdx.x]);
  const SweepModel& m = c.model();
  const int n = c.n, N = c.N, Q = c.Q;
  const Team tm = sweepTeam(c.tid);
  const int slot = c.tid / kTeam;
  const bool teamOn = slot < c.nl;
  const int myLocus = c.b.firstLocus + slot;
  unsigned int none[2] = {0u, 0u};
  if (teamOn) {
    TreeView t = sweepTreeView(c, d, slot);
    double* coal = c.coal(slot);
    uint8_t* np = c.pop(slot);
    // ---- the previous global move of this locus: k_smp_global_resolve
    const SmpProposal prev = sd.prop[myLocus];
    if (gm.prevKind >= 0 && prev.valid) {
      if (gm.prevAccept) {
        for (int x = tm.j; x < N; x += kTeam) commitNode(t, x);
        for (int p = tm.j; p < Q; p += kTeam) coal[p] = gm.prevKind == 0 ? sd.coalT[(size_t)myLocus * Q + p] : coal[p] * gm.prevC;
        if (tm.j == 0) commitLocus(t);
      } else {
        for (int x = tm.j; x < N; x += kTeam) revertNode(t, x);
        if (tm.j == 0) revertLocus(t);
      }
    }
    __syncwarp(tm.mask);
    // ---- this move's proposal
    SmpProposal pr = smpNoProposal();
    if (gm.kind == 0) {   // k_smp_tau_propose
      const int A = gm.A;
      pr.pop = A;
      if (*t.root >= n) {
        // the model this launch sees still holds the old split time of A unless the previous move changed it: the host
        // uploads the model after every accepted move, so m.tau is current
        const bool isRoot = m.father[A] < 0;
        int s0 = -1, s1 = -1;   // A's sons (none for a current population: then its SAMPLE AGE moves, GPhoCS.c:4006)
        for (int p = 0; p < Q; p++)
          if (m.father[p] == A) { if (s0 < 0) s0 = p; else s1 = p; }
        int n0 = 0, n1 = 0;
        for (int x = tm.j; x < N; x += kTeam) {
          int which = 0;   // 1: lower band, 2: upper band, 3: sample of A
          const int q = np[x];
          const double a = t.age[x];
          if (x >= n) {
            if (q == A) { if (isRoot || (a > gm.tauOld && a < gm.ub)) which = 2; }
            else if ((q == s0 || q == s1) && a > gm.lb && a < gm.tauOld) which = 1;
          } else if (s0 < 0 && q == A) {
            which = 3;
          }
          if (which) {
            const double an = which == 3 ? gm.tauNew : (which == 1 || isRoot ? gm.lb + (a - gm.lb) * gm.f0 : gm.ub + (a - gm.ub) * gm.f1);
            adjustAge(t, x, an);
          }
          n0 += which == 1;
          n1 += which == 2;
        }
        n0 = teamSumI(tm, n0);
        n1 = teamSumI(tm, n1);
        __syncwarp(tm.mask);
        // statistics under the proposed split time -> the pending arrays in HBM (committed by the next launch if accepted)
        double* terms = reinterpret_cast<double*>(c.sched(slot));         // team scratch (the schedule is built later):
        double* coalT = terms + c.NI;                                      // NI + Q doubles fit NI 32-byte entries,
        int* ncoalT = reinterpret_cast<int*>(c.walk(slot));                // Q ints the walk words
        teamStatsAll(tm, m, t.age, np, n, N, A, gm.tauNew, terms, coalT, ncoalT);
        for (int p = tm.j; p < Q; p += kTeam) {
          sd.coalT[(size_t)myLocus * Q + p] = coalT[p];
          sd.ncoalT[(size_t)myLocus * Q + p] = ncoalT[p];
        }
        double delta = 0.0;   // smpGenDelta, by the leader (population order)
        if (tm.j == 0) {
          const int* ncoal = c.ncoal(slot);
          for (int p = 0; p < Q; p++) {
            delta -= (coalT[p] - coal[p]) / m.theta[p];
            if (ncoalT[p] != ncoal[p]) delta += (double)(ncoalT[p] - ncoal[p]) * log(2.0 / m.theta[p]);
          }
        }
        pr.genDelta = delta;
        pr.ntj0 = n0;
        pr.ntj1 = n1;
        pr.valid = 1;
        __syncwarp(tm.mask);
      }
    } else if (gm.kind == 1) {   // k_smp_scale_propose
      if (*t.root >= n) {
        for (int x = tm.j; x < N; x += kTeam) adjustAge(t, x, gm.c * t.age[x]);
        pr.valid = 1;
      }
    }
    if (tm.j == 0) c.prop()[slot] = pr;
    if (gm.kind >= 0) sweepMarkAndCompact(c, tm, slot);
  }
  if (gm.kind >= 0) sweepWalkAndRoot(c, d, sweepListSchedule(c));
  sweepWriteBack(c, d, sd, none, false, true);
}

}  // namespace gphocs
