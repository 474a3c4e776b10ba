// sampler_mig.cuh — device-resident MCMC update steps for population trees WITH migration bands.
//
// Adds to sampler_kernels.cuh what the reference does with its per-population event chains when lineages can
// migrate (patch.c): UpdateGB_MigrationNode (GPhoCS.c:2437-2590), the migration part of UpdateGB_MigSPR /
// traceLineage (patch.c:886-1331), UpdateMigRates (GPhoCS.c:3110-3210) and the migration-band handling of UpdateTau
// and mixing.  Representation: a genealogy branch is cut into SEGMENTS (population, from, to) at population
// boundaries and at its migration events; everything the genealogy likelihood needs follows from the segment list
// without ordering events:
//     coal_stats[p] = sum over pairs of segments in p of 2 * overlap         ( = integral of n(n-1) )
//     mig_stats[b]  = sum over segments in target(b) of overlap with the band's live interval  ( = integral of n )
// One warp per locus; the locus' nodes, migration events and segments live in that warp's shared memory.
#pragma once
#include "sampler_kernels.cuh"

namespace gphocs {

struct SmgWarp {      // per-warp shared-memory view (carved by smgCarve)
  double* age;        // [N]
  NodeRec* node;      // [N] topology + flag bytes (the proposal kernels edit this copy and write it back once)
  uint8_t* pop;       // [N]
  double* migAge;     // [kSmpMaxMigs]
  int16_t* migBranch; // [kSmpMaxMigs]
  uint8_t* migBand;   // [kSmpMaxMigs]
  int* numMigs;       // [1]
  int* segCount;      // [1]
  int* bad;           // [1] inconsistency flag raised while cutting branches into segments
  double* segT0;      // [S]
  double* segT1;      // [S]
  int16_t* segBranch; // [S]
  uint16_t* list;     // [S] segments of one population, compacted (smgPopStats)
  uint8_t* segPop;    // [S]
  double* coal;       // [Q]
  double* mig;        // [B]
  int* ncoal;         // [Q]
  int* nmig;          // [B]
  int maxSegs;
};

__host__ __device__ inline int smgMaxSegs(int N) { return 4 * N + 2 * kSmpMaxMigs + 8; }
__host__ __device__ inline size_t smgWarpBytes(int N, int Q, int B) {
  const size_t S = (size_t)smgMaxSegs(N);
  size_t b = (size_t)N * 16 + kSmpMaxMigs * 8 + S * 16 + (size_t)Q * 8 + (size_t)B * 8;  // doubles and node records first
  b += 3 * 4 + (size_t)Q * 4 + (size_t)B * 4;                                            // ints
  b += kSmpMaxMigs * 2 + S * 4;                                                          // int16
  b += (size_t)N + kSmpMaxMigs + S;                                                      // bytes
  return (b + 15) & ~(size_t)15;
}
__device__ inline SmgWarp smgCarve(unsigned char* base, int N, int Q, int B) {
  SmgWarp w;
  const int S = smgMaxSegs(N);
  w.maxSegs = S;
  double* dp = reinterpret_cast<double*>(base);
  w.age = dp; dp += N;
  w.node = reinterpret_cast<NodeRec*>(dp); dp += N;
  w.migAge = dp; dp += kSmpMaxMigs;
  w.segT0 = dp; dp += S;
  w.segT1 = dp; dp += S;
  w.coal = dp; dp += Q;
  w.mig = dp; dp += B;
  int* ip = reinterpret_cast<int*>(dp);
  w.numMigs = ip++; w.segCount = ip++; w.bad = ip++;
  w.ncoal = ip; ip += Q;
  w.nmig = ip; ip += B;
  int16_t* sp = reinterpret_cast<int16_t*>(ip);
  w.migBranch = sp; sp += kSmpMaxMigs;
  w.segBranch = sp; sp += S;
  w.list = reinterpret_cast<uint16_t*>(sp); sp += S;
  uint8_t* bp = reinterpret_cast<uint8_t*>(sp);
  w.pop = bp; bp += N;
  w.migBand = bp; bp += kSmpMaxMigs;
  w.segPop = bp;
  return w;
}

// a band is live while both populations exist; a current population exists from time 0 whatever the age of its
// samples (updateMigrationBandTimes uses pops[]->age, PopulationTree.c:448)
__device__ inline double smgPopBirth(const SmpModel& m, int p, int ovPop, double ovTau) { return p < m.C ? 0.0 : smpTau(m, p, ovPop, ovTau); }
__device__ inline double smgBandStart(const SmpModel& m, int b, int ovPop, double ovTau) {
  return fmax(smgPopBirth(m, m.bandSrc[b], ovPop, ovTau), smgPopBirth(m, m.bandTgt[b], ovPop, ovTau));
}
__device__ inline double smgBandEnd(const SmpModel& m, int b, int ovPop, double ovTau) {
  return fmin(smpPopEnd(m, m.bandSrc[b], ovPop, ovTau), smpPopEnd(m, m.bandTgt[b], ovPop, ovTau));
}

// the locus' genealogy and migration events: HBM -> this warp's shared memory
__device__ inline void smgLoad(const SmgWarp& w, const StoreDev& d, const SmpDev& sd, int l, int lane) {
  const int N = d.N;
  for (int x = lane; x < N; x += 32) {
    const size_t o = (size_t)l * N + x;
    w.age[x] = d.age[o];
    w.node[x] = d.node[o];
    w.pop[x] = sd.nodePop[o];
  }
  if (lane < kSmpMaxMigs) {
    const size_t o = (size_t)l * kSmpMaxMigs + lane;
    w.migAge[lane] = sd.migAge[o];
    w.migBranch[lane] = sd.migBranch[o];
    w.migBand[lane] = sd.migBand[o];
  }
  if (lane == 0) { *w.numMigs = sd.numMigs[l]; *w.bad = 0; }
  __syncwarp();
}

// migration events shared memory -> HBM (after a proposal rewired them)
__device__ inline void smgStoreMigs(const SmgWarp& w, const SmpDev& sd, int l, int lane) {
  __syncwarp();
  if (lane < kSmpMaxMigs) {
    const size_t o = (size_t)l * kSmpMaxMigs + lane;
    sd.migAge[o] = w.migAge[lane];
    sd.migBranch[o] = w.migBranch[lane];
    sd.migBand[o] = w.migBand[lane];
  }
  if (lane == 0) sd.numMigs[l] = *w.numMigs;
}
// current events -> saved copy (restored if the proposal is rejected); `w` still holds the events as smgLoad read them
__device__ inline void smgSaveMigs(const SmgWarp& w, const SmpDev& sd, int l, int lane) {
  if (lane < kSmpMaxMigs) {
    const size_t o = (size_t)l * kSmpMaxMigs + lane;
    sd.svMigAge[o] = w.migAge[lane];
    sd.svMigBranch[o] = w.migBranch[lane];
    sd.svMigBand[o] = w.migBand[lane];
  }
  if (lane == 0) sd.svNumMigs[l] = *w.numMigs;
}
// the same from the HBM copy, for kernels that do not stage the locus
__device__ inline void smgSaveMigs(const SmpDev& sd, int l, int lane) {
  if (lane < kSmpMaxMigs) {
    const size_t o = (size_t)l * kSmpMaxMigs + lane;
    sd.svMigAge[o] = sd.migAge[o];
    sd.svMigBranch[o] = sd.migBranch[o];
    sd.svMigBand[o] = sd.migBand[o];
  }
  if (lane == 0) sd.svNumMigs[l] = sd.numMigs[l];
}
__device__ inline void smgRestoreMigs(const SmpDev& sd, int l, int lane) {
  if (lane < kSmpMaxMigs) {
    const size_t o = (size_t)l * kSmpMaxMigs + lane;
    sd.migAge[o] = sd.svMigAge[o];
    sd.migBranch[o] = sd.svMigBranch[o];
    sd.migBand[o] = sd.svMigBand[o];
  }
  if (lane == 0) sd.numMigs[l] = sd.svNumMigs[l];
}
// saved copy -> the current events in HBM and their staged copy
__device__ inline void smgRestoreMigs(const SmgWarp& w, const SmpDev& sd, int l, int lane) {
  if (lane < kSmpMaxMigs) {
    const size_t o = (size_t)l * kSmpMaxMigs + lane;
    const double a = sd.svMigAge[o];
    const int16_t br = sd.svMigBranch[o];
    const uint8_t bd = sd.svMigBand[o];
    sd.migAge[o] = a; sd.migBranch[o] = br; sd.migBand[o] = bd;
    w.migAge[lane] = a; w.migBranch[lane] = br; w.migBand[lane] = bd;
  }
  if (lane == 0) { const int k = sd.svNumMigs[l]; sd.numMigs[l] = k; *w.numMigs = k; }
}

// The proposal kernels of the per-locus sweeps edit the genealogy in this warp's shared memory: node records and ages
// of the view are the staged copies (one coalesced load by smgLoad, one coalesced store by smgStoreTree), the saved
// copies, the root and the log-likelihoods stay where they are in HBM (written, hardly ever read).  A tree edit by one
// lane is then a handful of shared-memory accesses instead of a chain of dependent HBM round trips.
__device__ inline TreeView smgStagedView(const TreeView& t, const SmgWarp& w) {
  TreeView s = t;
  s.node = w.node;
  s.age = w.age;
  return s;
}
__device__ inline void smgStoreTree(const SmgWarp& w, const StoreDev& d, int l, int lane) {
  __syncwarp();
  const int N = d.N;
  for (int x = lane; x < N; x += 32) {
    const size_t o = (size_t)l * N + x;
    d.node[o] = w.node[x];
    d.age[o] = w.age[x];
  }
}

// Segment list of the genealogy.  Every lane walks one branch (from its node up to its father, or for ever above the
// root) through the populations it visits — up at population ends, sideways (target -> source) at its migration
// events — one segment per round; the lanes that still have a segment to report in a round take consecutive slots
// (one ballot), so every branch is walked exactly once and the list needs no prefix pass.  The order of the list is
// (round of 32 branches, step of the walk, branch); nothing depends on it but the order of the sums taken over it.
// skipBranch: a branch left out (the pruned lineage of an SPR, -1: none); relabelFrom/relabelTo: segments of branch
// relabelFrom are reported as belonging to relabelTo (the pruned father's upper branch continues its remaining
// child's lineage).  Raises *w.bad if a path is inconsistent — a migration event outside its band's target population
// or live interval, an event beyond its branch, a branch that does not end in the population of the father's
// coalescence — or if the list overflows.
__device__ inline void smgBuildSegments(const SmpModel& m, const SmgWarp& w, int N, int root, int lane, int ovPop, double ovTau,
                                        int skipBranch, int relabelFrom, int relabelTo) {
  const int nm = *w.numMigs;
  const unsigned below = (1u << lane) - 1u;
  int base = 0, bad = 0;
  for (int x0 = 0; x0 < N; x0 += 32) {
    const int x = x0 + lane;
    bool live = x < N && x != skipBranch;
    int pop = 0, fa = -1;
    double t = 0.0, tEnd = kSmpInf;
    unsigned used = 0, mine = 0;   // mine: the events that sit on this branch (most branches carry none: no scan per step then)
    if (live) {
      pop = w.pop[x];
      t = w.age[x];
      fa = w.node[x].father;
      if (fa >= 0) tEnd = w.age[fa];
      bad |= (fa >= 0 && tEnd < t) || (fa < 0 && x != root);
      for (int k = 0; k < nm; k++) mine |= (unsigned)(w.migBranch[k] == x) << k;
    }
    const int label = x == relabelFrom ? relabelTo : x;
    for (int it = 0; it < 2 * kSmpMaxPops + kSmpMaxMigs + 2; it++) {
      const unsigned ballot = __ballot_sync(0xffffffffu, live);
      if (!ballot) break;
      if (live) {
        int mi = -1;   // earliest migration event of this branch not yet passed
        double mAge = kSmpInf;
        for (unsigned rest = mine & ~used; rest; rest &= rest - 1) {
          const int k = __ffs(rest) - 1;
          if (w.migAge[k] < mAge) { mi = k; mAge = w.migAge[k]; }
        }
        const double popEnd = m.father[pop] >= 0 ? smpTau(m, m.father[pop], ovPop, ovTau) : kSmpInf;
        const double tNext = fmin(tEnd, fmin(popEnd, mAge));
        const int k = base + __popc(ballot & below);
        if (k < w.maxSegs) { w.segPop[k] = (uint8_t)pop; w.segT0[k] = t; w.segT1[k] = tNext; w.segBranch[k] = (int16_t)label; }
        if (mi >= 0 && mAge <= tEnd && mAge <= popEnd) {   // sideways: target -> source of the band
          const int b = w.migBand[mi];
          if (pop != m.bandTgt[b] || mAge < t || mAge < smgBandStart(m, b, ovPop, ovTau) || mAge > smgBandEnd(m, b, ovPop, ovTau)) bad = 1;
          pop = m.bandSrc[b];
          t = mAge;
          used |= 1u << mi;
        } else {
          if (mi >= 0 && tEnd <= popEnd) bad = 1;   // an event of this branch lies beyond the branch
          if (tEnd <= popEnd || m.father[pop] < 0) {   // reached the father / the root population never ends
            if (fa >= 0 && pop != w.pop[fa]) bad = 1;
            live = false;
          } else {
            pop = m.father[pop];
            t = tNext;
          }
        }
      }
      base += __popc(ballot);
    }
  }
  if (base > w.maxSegs) { bad = 1; base = w.maxSegs; }
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    *w.segCount = base;
    if (bad) *w.bad = 1;
  }
  __syncwarp();
}

// Statistics that a move inside population p can change: coal[p] and mig[b] of the bands whose target is p.  The
// segments of p are compacted into w.list and only they are paired.  Leaves the results in w.coal[p] / w.mig[b];
// ncoal / nmig are unchanged by such a move.
__device__ inline void smgPopStats(const SmpModel& m, const SmgWarp& w, int lane, int p, int ovPop = -1, double ovTau = 0.0) {
  const int S = *w.segCount;
  uint16_t* list = w.list;
  int count = 0;
  for (int i0 = 0; i0 < S; i0 += 32) {
    const int i = i0 + lane;
    const bool in = i < S && w.segPop[i] == p;
    const unsigned ballot = __ballot_sync(0xffffffffu, in);
    if (in) list[count + __popc(ballot & ((1u << lane) - 1u))] = (uint16_t)i;
    count += __popc(ballot);
  }
  __syncwarp();
  // every unordered pair once, the same number of partners for every lane: entry a meets the (count-1)/2 entries that
  // follow it around the circle and, for an even count, the entry opposite if a is in the first half
  double c = 0.0;
  const int half = (count - 1) >> 1;
  for (int a = lane; a < count; a += 32) {
    const int i = list[a];
    const double a0 = w.segT0[i], a1 = w.segT1[i];
    const int partners = half + (((count & 1) == 0 && a < (count >> 1)) ? 1 : 0);
    int bq = a;
    for (int k = 0; k < partners; k++) {
      bq = bq + 1 == count ? 0 : bq + 1;
      const int j = list[bq];
      const double ov = fmin(a1, w.segT1[j]) - fmax(a0, w.segT0[j]);
      if (ov > 0.0) c += ov;
    }
  }
  c = 2.0 * warpSumD(c);
  if (lane == 0) w.coal[p] = c;
  for (int b = 0; b < m.B; b++) {
    if (m.bandTgt[b] != p) continue;
    const double s0 = smgBandStart(m, b, ovPop, ovTau), s1 = smgBandEnd(m, b, ovPop, ovTau);
    double g = 0.0;
    for (int a = lane; a < count; a += 32) {
      const int i = list[a];
      const double ov = fmin(s1, w.segT1[i]) - fmax(s0, w.segT0[i]);
      if (ov > 0.0) g += ov;
    }
    g = warpSumD(g);
    if (lane == 0) w.mig[b] = g;
  }
  __syncwarp();
}

// statistics of the locus from its segments -> w.coal / w.ncoal / w.mig / w.nmig (deterministic summation order):
// population by population (pairing the segments of one population at a time instead of filtering all pairs)
__device__ inline void smgStats(const SmpModel& m, const SmgWarp& w, int n, int N, int lane, int ovPop, double ovTau) {
  const int Q = m.Q, B = m.B;
  for (int p = 0; p < Q; p++) smgPopStats(m, w, lane, p, ovPop, ovTau);
  for (int p = lane; p < Q; p += 32) {
    int k = 0;
    for (int x = n; x < N; x++) k += w.pop[x] == p;
    w.ncoal[p] = k;
  }
  for (int b = lane; b < B; b += 32) {
    int k = 0;
    for (int e = 0; e < *w.numMigs; e++) k += w.migBand[e] == b;
    w.nmig[b] = k;
  }
  __syncwarp();
}

// genealogy log-density from statistics (gtreeLnLikelihood, patch.c:2702-2723)
__device__ inline double smgLnL(const SmpModel& m, const double* coal, const int* ncoal, const double* mig, const int* nmig) {
  double v = 0.0;
  for (int p = 0; p < m.Q; p++) v += (double)ncoal[p] * log(2.0 / m.theta[p]) - coal[p] / m.theta[p];
  for (int b = 0; b < m.B; b++)
    if (m.migRate[b] > 0.0) v += (double)nmig[b] * log(m.migRate[b]) - mig[b] * m.migRate[b];
  return v;
}

// w.* statistics -> the locus' stored (pending = 0) or pending (1) arrays
__device__ inline void smgWriteStats(const SmpModel& m, const SmgWarp& w, const SmpDev& sd, int l, int lane, int pending) {
  for (int p = lane; p < m.Q; p += 32) {
    (pending ? sd.coalT : sd.coal)[(size_t)l * m.Q + p] = w.coal[p];
    (pending ? sd.ncoalT : sd.ncoal)[(size_t)l * m.Q + p] = w.ncoal[p];
  }
  for (int b = lane; b < m.B; b += 32) {
    (pending ? sd.migT : sd.mig)[(size_t)l * m.B + b] = w.mig[b];
    (pending ? sd.nmigT : sd.nmig)[(size_t)l * m.B + b] = w.nmig[b];
  }
}
// the locus' stored statistics -> w.*
__device__ inline void smgLoadStats(const SmpModel& m, const SmgWarp& w, const SmpDev& sd, int l, int lane) {
  for (int p = lane; p < m.Q; p += 32) { w.coal[p] = sd.coal[(size_t)l * m.Q + p]; w.ncoal[p] = sd.ncoal[(size_t)l * m.Q + p]; }
  for (int b = lane; b < m.B; b += 32) { w.mig[b] = sd.mig[(size_t)l * m.B + b]; w.nmig[b] = sd.nmig[(size_t)l * m.B + b]; }
  __syncwarp();
}
__device__ inline double smgStoredLnL(const SmpModel& m, const SmpDev& sd, int l) {
  return smgLnL(m, sd.coal + (size_t)l * m.Q, sd.ncoal + (size_t)l * m.Q, sd.mig + (size_t)l * m.B, sd.nmig + (size_t)l * m.B);
}

#define SMG_PROLOGUE                                                                          \
  extern __shared__ __align__(16) unsigned char smgSmem[];                                    \
  SMP_STAGE_MODEL                                                                             \
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;                                  \
  const int l = sd.l0 + blockIdx.x * kSmpLociPerCta + wid;                                    \
  if (l >= sd.l1) return;                                                                     \
  const SmpModel& m = smpModelShared;                                                         \
  const int n = d.n, N = d.N;                                                                 \
  const SmgWarp w = smgCarve(smgSmem + (size_t)wid * smgWarpBytes(N, m.Q, m.B), N, m.Q, m.B); \
  const TreeView t = deviceView(d, l);                                                        \
  (void)n; (void)lane;

// ------------------------------------------------------------------------------------------ statistics from scratch
__global__ void __launch_bounds__(kSmpThreads) k_smg_stats(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int pending,
                                                           int* __restrict__ bad) {
  SMG_PROLOGUE
  if (*t.root < n) return;
  smgLoad(w, d, sd, l, lane);
  smgBuildSegments(m, w, N, *t.root, lane, -1, 0.0, -1, -1, -1);
  smgStats(m, w, n, N, lane, -1, 0.0);
  smgWriteStats(m, w, sd, l, lane, pending);
  if (bad && lane == 0) bad[l] += *w.bad;
}

__device__ inline void smgResolve(const StoreDev& d, const SmpDev& sd, const SmpModel& m, const TreeView& t, const SmgWarp* w, int l,
                                  int lane, int N, int kind, unsigned long long seed, unsigned long long step);

// ------------------------------------------------------------------------------------------ coalescence-time move
// UpdateGB_InternalNode with migration: the node stays in its population and between the events next to it on
// the three branches it touches (GPhoCS.c:2316-2351, findFirstMig / findLastMig patch.c:374-410).
__device__ inline void smgAgeProposeBody(const StoreDev& d, const SmpDev& sd, const SmpModel& m, const SmgWarp& w, const TreeView& t,
                                         int l, int lane, int n, int N, int inode, double finetune, unsigned long long seed,
                                         unsigned long long step) {
  SmpProposal pr = smpNoProposal();
  pr.node = inode;
  const int root = *t.root;
  if (root < n) { if (lane == 0) sd.prop[l] = pr; return; }
  double tnew = 0.0;
  int valid = 0;
  if (lane == 0) {
    const int pop = w.pop[inode];
    const double told = w.age[inode];
    const NodeRec rec = t.node[inode];
    double lo = m.tau[pop];
    double hi = m.father[pop] >= 0 ? m.tau[m.father[pop]] : kOldAge;
    double up = inode != root ? w.age[rec.father] : kSmpInf;   // first event above on the node's own branch
    double dl = w.age[rec.left], dr = w.age[rec.right];         // last events below on the children's branches
    for (int k = 0; k < *w.numMigs; k++) {
      const int br = w.migBranch[k];
      if (br == inode) up = fmin(up, w.migAge[k]);
      if (br == rec.left) dl = fmax(dl, w.migAge[k]);
      if (br == rec.right) dr = fmax(dr, w.migAge[k]);
    }
    lo = fmax(lo, fmax(dl, dr));
    hi = fmin(hi, up);
    SmpRng rng(seed, (unsigned long long)l, step);
    tnew = smpReflect(told + finetune * rng.normal2(), lo, hi);
    valid = fabs(tnew - told) >= 1e-15;
    if (valid) adjustAge(t, inode, tnew);   // t.age is w.age: the staged copy takes the new age
  }
  valid = __shfl_sync(0xffffffffu, valid, 0);
  if (valid) {
    __syncwarp();
    const int p = w.pop[inode];
    smgBuildSegments(m, w, N, root, lane, -1, 0.0, -1, -1, -1);
    smgPopStats(m, w, lane, p);
    if (lane == 0) {
      // only population p's coal statistic and the bands into p change (the node stays in p)
      double delta = -(w.coal[p] - sd.coal[(size_t)l * m.Q + p]) / m.theta[p];
      sd.coalT[(size_t)l * m.Q + p] = w.coal[p];
      for (int b = 0; b < m.B; b++)
        if (m.bandTgt[b] == p) {
          sd.migT[(size_t)l * m.B + b] = w.mig[b];
          if (m.migRate[b] > 0.0) delta -= (w.mig[b] - sd.mig[(size_t)l * m.B + b]) * m.migRate[b];
        }
      pr.genDelta = delta;
      pr.pop = p;
      pr.valid = *w.bad ? 0 : 1;
      if (*w.bad) { revertNode(t, inode); }   // cannot happen for a move inside its bounds; stay safe
    }
  }
  if (lane == 0) sd.prop[l] = pr;
}
// The proposal kernels of a sweep: stage the locus, settle the previous proposal of the sweep on the staged copy,
// propose, write the genealogy back.
__global__ void __launch_bounds__(kSmpThreads)
k_smg_age_propose(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int inode, double finetune, unsigned long long seed,
                  unsigned long long step, int pendKind, unsigned long long pendStep) {
  SMG_PROLOGUE
  smgLoad(w, d, sd, l, lane);
  const TreeView ts = smgStagedView(t, w);
  if (pendKind >= 0) smgResolve(d, sd, m, ts, &w, l, lane, N, pendKind, seed, pendStep);   // the previous proposal of this locus
  smgAgeProposeBody(d, sd, m, w, ts, l, lane, n, N, inode, finetune, seed, step);
  smgStoreTree(w, d, l, lane);
}

// ------------------------------------------------------------------------------------------ migration-time moves
// UpdateGB_MigrationNode (GPhoCS.c:2437-2590): every migration event of the locus in turn moves inside its band's
// live interval and between the events next to it on its branch; the data likelihood is not involved, so the
// whole sweep of a locus (propose, statistics, accept / reject) is done here, one launch for all loci.
__global__ void __launch_bounds__(kSmpThreads)
k_smg_mignode_sweep(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, double finetune, unsigned long long seed,
                    unsigned long long step) {
  SMG_PROLOGUE
  const int root = *t.root;
  if (root < n) return;
  smgLoad(w, d, sd, l, lane);
  const int nm = *w.numMigs;
  unsigned long long acc = 0, tried = 0;
  double cur = smgStoredLnL(m, sd, l);
  smgLoadStats(m, w, sd, l, lane);
  for (int k = 0; k < nm; k++) {
    double told = 0.0, tnew = 0.0;
    int valid = 0;
    if (lane == 0) {
      const int br = w.migBranch[k], b = w.migBand[k];
      told = w.migAge[k];
      double lo = fmax(smgBandStart(m, b, -1, 0.0), w.age[br]);
      double hi = fmin(smgBandEnd(m, b, -1, 0.0), w.node[br].father >= 0 ? w.age[w.node[br].father] : kOldAge);
      for (int j = 0; j < nm; j++)
        if (j != k && w.migBranch[j] == br) {
          if (w.migAge[j] < told) lo = fmax(lo, w.migAge[j]);
          else hi = fmin(hi, w.migAge[j]);
        }
      SmpRng rng(seed, (unsigned long long)l, step * 16ull + (unsigned long long)k);
      tnew = smpReflect(told + finetune * rng.normal2(), lo, hi);
      valid = fabs(tnew - told) >= 1e-15;
      if (valid) w.migAge[k] = tnew;
    }
    valid = __shfl_sync(0xffffffffu, valid, 0);
    tried++;
    if (!valid) { acc++; continue; }
    told = __shfl_sync(0xffffffffu, told, 0);
    __syncwarp();
    smgBuildSegments(m, w, N, root, lane, -1, 0.0, -1, -1, -1);
    // the event ends a segment in its band's target population and begins one in the source population: only the
    // statistics of these two populations and of the bands into them can change
    const int band = w.migBand[k];
    smgPopStats(m, w, lane, m.bandTgt[band]);
    smgPopStats(m, w, lane, m.bandSrc[band]);
    int ok = 0;
    if (lane == 0) {
      const double next = smgLnL(m, w.coal, w.ncoal, w.mig, w.nmig);
      const double lnacc = next - cur;
      ok = !*w.bad && lnacc >= 0.0;
      if (!ok && !*w.bad) {
        SmpRng rng(seed, (unsigned long long)l, step * 16ull + (unsigned long long)k + 8ull);
        ok = rng.uniform() < exp(lnacc);
      }
      if (ok) cur = next;
      else { w.migAge[k] = told; *w.bad = 0; }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    cur = __shfl_sync(0xffffffffu, cur, 0);
    __syncwarp();   // lane 0 has read the proposed statistics
    if (ok) { smgWriteStats(m, w, sd, l, lane, 0); acc++; }
    else smgLoadStats(m, w, sd, l, lane);   // rejected (one proposal in ten): back to the stored statistics
    __syncwarp();
  }
  smgStoreMigs(w, sd, l, lane);
  if (lane == 0 && tried) { atomicAdd(sd.accepted + 5, acc); atomicAdd(sd.accepted + 6, tried); }
}

// ------------------------------------------------------------------------------------------ subtree prune and regraft
// traceLineage with migration (patch.c:886-1331): the pruned lineage is re-simulated from the coalescent with
// migration conditional on the rest of the genealogy.  While it sits in population p the competing clocks are a
// coalescence with every other lineage segment in p (rate 2/theta_p while both are there) and a migration
// through every live band into p (rate m_b); the earliest ring over the warp decides: coalescence ends the
// walk, a migration moves the lineage to the band's source population, no ring before the population ends moves
// it to the parent population.  Proposal = conditional prior, so acceptance is the data-likelihood ratio alone
// (GPhoCS.c:2702-2706); more than MAX_MIGS events in the genealogy make the proposal invalid (res < 0, :2706).
__device__ inline void smgSprProposeBody(const StoreDev& d, const SmpDev& sd, const SmpModel& m, const SmgWarp& w, const TreeView& t,
                                         int l, int lane, int n, int N, int node, unsigned long long seed, unsigned long long step) {
  SmpProposal pr = smpNoProposal();
  const int root = *t.root;
  if (root < n || node == root) { if (lane == 0) sd.prop[l] = pr; return; }
  const int F = w.node[node].father;
  const NodeRec recF = t.node[F];
  const int S = recF.left + recF.right - node;
  // pruned genealogy: without the branch of `node`; the father's upper branch continues the sibling's lineage
  smgBuildSegments(m, w, N, root, lane, -1, 0.0, node, F, S);
  const int nSeg = *w.segCount;
  double now = w.age[node];
  int pop = w.pop[node];
  int newBand[kSmpMaxMigs];
  double newAge[kSmpMaxMigs];
  int numNew = 0, target = -1, fail = *w.bad;
  int kept = 0;   // events of the genealogy that survive the pruning
  for (int k = 0; k < *w.numMigs; k++) kept += w.migBranch[k] != node;
  // one clock per (sojourn, segment) and per (sojourn, band): draws of one family of streams under (locus, step)
  const SmpRng rng(seed, (unsigned long long)l, step);
  for (int it = 0; it < 4 * kSmpMaxPops + 2 * kSmpMaxMigs && target < 0 && !fail; it++) {
    const double popEnd = m.father[pop] >= 0 ? m.tau[m.father[pop]] : kSmpInf;
    double bestT = kSmpInf;
    int bestKind = -1, bestId = -1;
    for (int i = lane; i < nSeg; i += 32) {
      if (w.segPop[i] != pop) continue;
      const double a = fmax(now, w.segT0[i]), b = fmin(popEnd, w.segT1[i]);
      if (b <= a) continue;
      const double T = a + rng.exponentialAt((unsigned long long)it * 8192ull + (unsigned long long)i) * m.theta[pop] * 0.5;
      if (T < b && T < bestT) { bestT = T; bestKind = 0; bestId = i; }
    }
    for (int b = lane; b < m.B; b += 32) {
      if (m.bandTgt[b] != pop || !(m.migRate[b] > 0.0)) continue;
      const double a = fmax(now, smgBandStart(m, b, -1, 0.0)), e = fmin(popEnd, smgBandEnd(m, b, -1, 0.0));
      if (e <= a) continue;
      const double T = a + rng.exponentialAt((unsigned long long)it * 8192ull + 4096ull + (unsigned long long)b) / m.migRate[b];
      if (T < e && T < bestT) { bestT = T; bestKind = 1; bestId = b; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double oT = __shfl_xor_sync(0xffffffffu, bestT, off);
      const int oK = __shfl_xor_sync(0xffffffffu, bestKind, off);
      const int oI = __shfl_xor_sync(0xffffffffu, bestId, off);
      if (oT < bestT || (oT == bestT && oK >= 0 && (bestKind < 0 || oK * 4096 + oI < bestKind * 4096 + bestId))) {
        bestT = oT; bestKind = oK; bestId = oI;
      }
    }
    if (bestKind == 0) {
      target = w.segBranch[bestId];
      now = bestT;
    } else if (bestKind == 1) {
      if (kept + numNew >= kSmpMaxMigs) { fail = 1; break; }
      newBand[numNew] = bestId;
      newAge[numNew] = bestT;
      numNew++;
      now = bestT;
      pop = m.bandSrc[bestId];
    } else {
      if (m.father[pop] < 0) { fail = 1; break; }
      now = popEnd;
      pop = m.father[pop];
    }
  }
  if (target >= 0 && !fail) {
    smgSaveMigs(w, sd, l, lane);
    __syncwarp();
    if (lane == 0) {
      // events of the genealogy after the move: the old lineage's go, the pruned father's upper branch joins the
      // sibling, events of the target branch above the new node move to the new father's upper branch, the
      // simulated ones sit on the regrafted branch
      int k2 = 0;
      const int nmOld = *w.numMigs;
      for (int k = 0; k < nmOld; k++) {
        int br = w.migBranch[k];
        if (br == node) continue;
        if (br == F) br = S;
        if (br == target && w.migAge[k] > now) br = F;
        w.migBranch[k2] = (int16_t)br; w.migBand[k2] = w.migBand[k]; w.migAge[k2] = w.migAge[k]; k2++;
      }
      for (int k = 0; k < numNew; k++) {
        w.migBranch[k2] = (int16_t)node; w.migBand[k2] = (uint8_t)newBand[k]; w.migAge[k2] = newAge[k]; k2++;
      }
      *w.numMigs = k2;
      uint8_t* np = sd.nodePop + (size_t)l * N;
      pr.pop = np[F];
      pr.node = F;
      spr(t, node, target, now);
      np[F] = (uint8_t)pop;
      pr.valid = 1;
      pr.ntj0 = 1;   // the migration events were rewired: restore them on rejection
    }
    smgStoreMigs(w, sd, l, lane);
  }
  if (lane == 0) sd.prop[l] = pr;
}
__global__ void __launch_bounds__(kSmpThreads, 8)
k_smg_spr_propose(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int node, unsigned long long seed,
                  unsigned long long step, int pendKind, unsigned long long pendStep) {
  SMG_PROLOGUE
  smgLoad(w, d, sd, l, lane);
  const TreeView ts = smgStagedView(t, w);
  if (pendKind >= 0) smgResolve(d, sd, m, ts, &w, l, lane, N, pendKind, seed, pendStep);
  smgSprProposeBody(d, sd, m, w, ts, l, lane, n, N, node, seed, step);
  smgStoreTree(w, d, l, lane);
}

// ------------------------------------------------------------------------------------------ split-time move
// Rubber band with migration: coalescences AND migration events located in the affected populations are rescaled
// (patch.c:699-712, 727, 741-748); the move is invalid (mig_conflict, GPhoCS.c:3545-3697) if afterwards some event
// lies outside its band's new live interval or out of order on its branch — found by re-cutting the branches.
__global__ void __launch_bounds__(kSmpThreads)
k_smg_tau_propose(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int A, double tauOld, double tauNew, double lb, double ub,
                  double f0, double f1) {
  SMG_PROLOGUE
  SmpProposal pr = smpNoProposal();
  pr.pop = A;
  const int root = *t.root;
  if (root < n) { if (lane == 0) sd.prop[l] = pr; return; }
  smgLoad(w, d, sd, l, lane);
  smgSaveMigs(w, sd, l, lane);
  const TreeView ts = smgStagedView(t, w);
  const bool isRoot = A == m.rootPop;
  // A current population: its SAMPLE AGE moves (UpdateSampleAge, GPhoCS.c:4006-4590): the leaves of A take the new age
  const int s0 = A >= m.C ? m.son0[A] : -1, s1 = A >= m.C ? m.son1[A] : -1;
  int n0 = 0, n1 = 0;
  for (int x0 = 0; x0 < N; x0 += 32) {
    const int x = x0 + lane;
    int which = 0;
    if (x < N) {
      const int q = w.pop[x];
      const double a = w.age[x];
      if (x >= n) {
        if (q == A) {
          if (isRoot || (a > tauOld && a < ub)) which = 2;
          else if (A < m.C && a < tauOld) which = 1;   // migrants that entered A before its samples: scaled towards 0
        } else if ((q == s0 || q == s1) && a > lb && a < tauOld) which = 1;
      } else if (A < m.C && q == A) {
        which = 3;
      }
      if (which) {
        const double an = which == 3 ? tauNew : (which == 1 || isRoot ? lb + (a - lb) * f0 : ub + (a - ub) * f1);
        adjustAge(ts, x, an);   // ts.age is w.age
      }
    }
    n0 += __popc(__ballot_sync(0xffffffffu, which == 1));
    n1 += __popc(__ballot_sync(0xffffffffu, which == 2));
  }
  {
    int which = 0;
    if (lane < *w.numMigs) {
      const int b = w.migBand[lane];
      const int ps = m.bandSrc[b], pt = m.bandTgt[b];
      const double a = w.migAge[lane];
      if (ps == A || pt == A) {
        if (isRoot || (a > tauOld && a < ub)) which = 2;
        else if (A < m.C && a < tauOld) which = 1;
      } else if ((ps == s0 || ps == s1 || pt == s0 || pt == s1) && a > lb && a < tauOld) which = 1;
      if (which) w.migAge[lane] = which == 1 || isRoot ? lb + (a - lb) * f0 : ub + (a - ub) * f1;
    }
    n0 += __popc(__ballot_sync(0xffffffffu, which == 1));
    n1 += __popc(__ballot_sync(0xffffffffu, which == 2));
  }
  __syncwarp();
  smgBuildSegments(m, w, N, root, lane, A, tauNew, -1, -1, -1);
  // The rubber band moves events of A and its two sons only, and only bands that touch one of the three change their
  // live interval: statistics of every other population and band keep their stored values.  (Pairing the segments of
  // one population at a time — a few of the ~70 — instead of all of them at once is what makes this kernel 2x faster.)
  smgLoadStats(m, w, sd, l, lane);
  unsigned long long touched = 1ull << A;
  if (s0 >= 0) touched |= (1ull << s0) | (1ull << s1);
  // + both ends of every band that touches one of them: its live interval moves, and its rescaled migration events
  // end a segment in the target population and begin one in the source population
  unsigned long long redo = touched;
  for (int b = 0; b < m.B; b++)
    if (((touched >> m.bandSrc[b]) | (touched >> m.bandTgt[b])) & 1ull) redo |= (1ull << m.bandTgt[b]) | (1ull << m.bandSrc[b]);
  for (unsigned long long rest = redo; rest; rest &= rest - 1) smgPopStats(m, w, lane, __ffsll((long long)rest) - 1, A, tauNew);
  smgWriteStats(m, w, sd, l, lane, 1);
  smgStoreMigs(w, sd, l, lane);
  smgStoreTree(w, d, l, lane);
  if (lane == 0) {
    // log-density of the pending state under the proposed split time: theta and rates are unchanged
    pr.genDelta = smgLnL(m, w.coal, w.ncoal, w.mig, w.nmig) - smgStoredLnL(m, sd, l);
    pr.ntj0 = n0;
    pr.ntj1 = n1;
    pr.valid = 1;
    pr.node = *w.bad ? -2 : -1;   // -2: migration conflict in this locus
    sd.prop[l] = pr;
  }
}

// joint rescaling of every node age and migration time by c (mixing, GPhoCS.c:4793-4801, 4822-4826)
__global__ void __launch_bounds__(kSmpThreads) k_smg_scale_propose(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, double c) {
  SMG_PROLOGUE
  (void)w;
  SmpProposal pr = smpNoProposal();
  if (*t.root >= n) {
    smgSaveMigs(sd, l, lane);
    for (int x = lane; x < N; x += 32) adjustAge(t, x, c * t.age[x]);
    if (lane < sd.numMigs[l]) sd.migAge[(size_t)l * kSmpMaxMigs + lane] *= c;
    pr.valid = 1;
  }
  if (lane == 0) sd.prop[l] = pr;
}

// per-locus accept / reject for models with migration (kind as in k_smp_accept).  t: the genealogy to settle — the HBM
// view, or the staged view of a proposal kernel (w != NULL: what a rejection restores goes to the staged copy as well)
__device__ inline void smgResolve(const StoreDev& d, const SmpDev& sd, const SmpModel& m, const TreeView& t, const SmgWarp* w, int l,
                                  int lane, int N, int kind, unsigned long long seed, unsigned long long step) {
  const SmpProposal pr = sd.prop[l];
  int ok = 0;
  if (pr.valid) {
    if (lane == 0) {
      const double lnacc = (*t.lnL - *t.savedLnL) + pr.genDelta;
      ok = lnacc >= 0.0;
      if (!ok) {
        SmpRng rng(seed, (unsigned long long)l, step);
        ok = rng.uniform() < exp(lnacc);
      }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (ok) {
      for (int x = lane; x < N; x += 32) commitNode(t, x);
      if (kind == 0) {   // the changed statistics of the proposed state were left pending by the proposal kernel
        if (lane == 0) sd.coal[(size_t)l * m.Q + pr.pop] = sd.coalT[(size_t)l * m.Q + pr.pop];
        for (int b = lane; b < m.B; b += 32)
          if (m.bandTgt[b] == pr.pop) sd.mig[(size_t)l * m.B + b] = sd.migT[(size_t)l * m.B + b];
      }
      if (lane == 0) commitLocus(t);
    } else {
      for (int x = lane; x < N; x += 32)
        if (t.node[x].flags & (F_RECALC | F_SAVED)) revertNode(t, x);   // a no-op on unmarked nodes
      if (kind == 1) {
        if (w) smgRestoreMigs(*w, sd, l, lane);
        else smgRestoreMigs(sd, l, lane);
      }
      if (lane == 0) {
        revertLocus(t);
        if (kind == 1) {
          sd.nodePop[(size_t)l * N + pr.node] = (uint8_t)pr.pop;
          if (w) w->pop[pr.node] = (uint8_t)pr.pop;
        }
      }
    }
  } else if (kind == 0) {
    ok = 1;
  }
  if (lane == 0 && ok) atomicAdd(sd.accepted + kind, 1ull);
  __syncwarp();
}
__global__ void __launch_bounds__(kSmpThreads)
k_smg_accept(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int kind, unsigned long long seed, unsigned long long step) {
  SMG_PROLOGUE
  (void)w;
  smgResolve(d, sd, m, t, nullptr, l, lane, N, kind, seed, step);
}

// one global accept / reject for every locus; how: 0 tau move (statistics <- pending), 1 rescaling (statistics *= c)
__global__ void __launch_bounds__(kSmpThreads) k_smg_global_resolve(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int accept,
                                                                    int how, double c) {
  SMG_PROLOGUE
  (void)w;
  if (!sd.prop[l].valid) return;
  if (accept) {
    for (int x = lane; x < N; x += 32) commitNode(t, x);
    for (int p = lane; p < m.Q; p += 32) {
      double* c0 = sd.coal + (size_t)l * m.Q + p;
      *c0 = how == 0 ? sd.coalT[(size_t)l * m.Q + p] : *c0 * c;
    }
    for (int b = lane; b < m.B; b += 32) {
      double* c0 = sd.mig + (size_t)l * m.B + b;
      *c0 = how == 0 ? sd.migT[(size_t)l * m.B + b] : *c0 * c;
    }
    if (lane == 0) commitLocus(t);
  } else {
    for (int x = lane; x < N; x += 32) revertNode(t, x);
    smgRestoreMigs(sd, l, lane);
    if (lane == 0) revertLocus(t);
  }
}

}  // namespace gphocs
