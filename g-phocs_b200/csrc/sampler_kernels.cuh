// sampler_kernels.cuh — device-resident MCMC update steps (proposals on the device).
//
// What the reference does on the host, one locus at a time, with a device round trip per proposal
//   UpdateGB_InternalNode  GPhoCS.c:2287-2424   coalescence-time moves
//   UpdateGB_MigSPR        GPhoCS.c:2598-2760   subtree prune + re-coalescence drawn from the coalescent prior
//   UpdateTheta            GPhoCS.c:3035-3103   multiplicative theta moves from the total statistics
//   UpdateTau              GPhoCS.c:3224-3990   split-time moves with the rubber band (patch.c:596-801)
//   mixing                 GPhoCS.c:4688-4900   joint rescaling of all times and thetas
// is done here for ALL loci per launch: one warp per locus proposes (its own counter-based random stream) and
// edits the device-resident genealogy through the same edit protocol the host API uses (tree_ops.cuh); k_eval
// (clv_kernels.cuh) evaluates every locus incrementally; the same warp accepts or rejects.  No host round
// trip inside a sweep.  Scope of this version: population trees WITHOUT migration bands, samples of age 0,
// constant locus rates — the model of BASELINE.json configs[1].  The target density is the reference's:
//   P(X|G) * prod_pops (2/theta)^ncoal exp(-coal_stats/theta) * Gamma priors on theta and tau
// (gtreeLnLikelihood patch.c:2702-2723; priors as used in GPhoCS.c:3063-3065, 3444-3447, 4741-4760).
// The chains are not bit-identical to the reference's (different random streams): parity is statistical
// (tests/test_gpu_sampler.py: prior recovery with uninformative data, posterior means against the reference chain).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "clv_kernels.cuh"

namespace gphocs {

constexpr int kSmpMaxPops = 39;     // 2*NSPECIES-1 (patch.h:19)
constexpr double kOldAge = 999.0;   // OLDAGE (patch.h:22): end of the root population
constexpr int kSmpMaxBands = 32;    // migration bands the device-resident steps handle (reference: MAX_MIG_BANDS 100)
constexpr int kSmpMaxMigs = 10;     // MAX_MIGS (patch.h:18): migration events per genealogy

struct SmpModel {
  int Q, C, rootPop, n;
  int father[kSmpMaxPops], son0[kSmpMaxPops], son1[kSmpMaxPops], samplesPerPop[kSmpMaxPops];
  int postOrder[kSmpMaxPops];
  int leavesBelow[kSmpMaxPops];       // haploid samples in the current populations under each population
  unsigned long long below[kSmpMaxPops];  // bit q: population q is this population or lies below it
  double theta[kSmpMaxPops], tau[kSmpMaxPops];
  double coalRate[kSmpMaxPops];       // 2 / theta, refreshed with every upload of the model (smpUploadModel)
  // migration bands (MigrationBand, PopulationTree.h:60-70): backwards in time a lineage in the target population
  // moves to the source population at rate migRate while both populations exist
  int B;
  int bandSrc[kSmpMaxBands], bandTgt[kSmpMaxBands];
  double migRate[kSmpMaxBands];
};

struct SmpProposal {   // what a proposal kernel leaves for the accept kernel, per locus
  double genDelta;     // change of the genealogy log-density
  double aux;          // new coal statistic of `pop` (age move)
  int pop;             // population whose statistic changes (age move) / previous population of the moved father (SPR)
  int node;            // moved father (SPR)
  int valid;           // 0: nothing was proposed for this locus
  int ntj0, ntj1;      // nodes moved by the lower / upper rubber band (tau move)
};

struct SmpDev {
  int L, Q;
  int l0, l1;          // loci [l0, l1) a warp-per-locus launch works on (a sweep may be split over streams)
  uint8_t* nodePop;    // [L][N] population of every genealogy node (nodePops, patch.h:123)
  double* coal;        // [L][Q] coal_stats per locus
  double* coalT;       // [L][Q] statistics of the pending proposal
  int* ncoal;          // [L][Q] num_coals per locus
  int* ncoalT;         // [L][Q] coalescence counts of the pending proposal
  SmpProposal* prop;   // [L]
  // migration events of every genealogy (genetree_migs, patch.h:138-148) and their saved copies
  int *numMigs, *svNumMigs;          // [L]
  int16_t *migBranch, *svMigBranch;  // [L][kSmpMaxMigs] genealogy branch (node below) carrying the event
  uint8_t *migBand, *svMigBand;      // [L][kSmpMaxMigs]
  double *migAge, *svMigAge;         // [L][kSmpMaxMigs]
  double *mig, *migT;                // [L][B] mig_stats per locus, stored / pending
  int *nmig, *nmigT;                 // [L][B] num_migs per locus
  unsigned long long* accepted;  // [8] acceptance counters per move kind
  double* partial;     // [blocks][kSmpPartials] block partial sums
};
constexpr int kSmpPartials = 5 + 2 * kSmpMaxPops + 2 * kSmpMaxBands;
constexpr int kSmpThreads = 128;

// Counter-based random numbers: draw k of stream (seed, locus, step) is a SplitMix64-style hash of its coordinates,
// so kernels need no generator state and every (locus, step) owns an independent stream whatever the launch order.
struct SmpRng {
  unsigned long long key, ctr;
  __device__ SmpRng(unsigned long long seed, unsigned long long stream, unsigned long long step) {
    key = mix(seed ^ mix(stream * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull) ^ mix(step * 0xBF58476D1CE4E5B9ull + 0x94D049BB133111EBull));
    ctr = 0;
  }
  __device__ static unsigned long long mix(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  __device__ unsigned long long next() { return mix(key + (++ctr) * 0x9E3779B97F4A7C15ull); }
  __device__ double uniform() { return ((double)(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }   // (0, 1)
  // standard normal by Box-Muller in single precision: proposals only need a symmetric, well-spread kernel
  __device__ double normal() {
    const unsigned long long r = next();
    const float u1 = ((float)(unsigned)(r >> 40) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((float)(unsigned)((r >> 8) & 0xffffffu) + 0.5f) * (1.0f / 16777216.0f);
    return (double)(sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831853f * u2));
  }
  // rnd2normal8 (utils.c:482-488): mixture of N(-m, s^2) and N(m, s^2) with m^2 + s^2 = 1, m^2/s^2 = 8
  __device__ double normal2() {
    const double z = 0.94280904158206336 + normal() * (1.0 / 3.0);
    return uniform() < 0.5 ? z : -z;
  }
  __device__ double exponential() { return -log(uniform()); }
  // draw `idx` of a family of independent streams under this (seed, stream, step): one hash per draw, for loops that
  // need one exponential per branch or segment (the competing clocks of the SPR proposal)
  // The clocks of a proposal only have to be exponential to proposal accuracy: the logarithm is taken in single
  // precision (relative error 1e-7 of the waiting time), through log1p of u - 1 above one half so that short waiting
  // times keep that relative accuracy; the uniform itself has its 53 bits.
  __device__ double exponentialAt(unsigned long long idx) const {
    const unsigned long long r = mix(key + (idx + 0x1000ull) * 0x9E3779B97F4A7C15ull);
    const double u = ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    return u > 0.5 ? -(double)log1pf((float)(u - 1.0)) : -(double)logf((float)u);
  }
};

// reflect (utils.c:333-398): folds x into (a, b)
__device__ inline double smpReflect(double x, double a, double b) {
  const double slack = 0.000000001;
  a += slack;
  b -= slack;
  if (b <= a) return (a + b) / 2.0;
  if (x < b && x > a) return x;
  double xn = x;
  if (xn <= a) xn = 2.0 * a - xn;
  const double dbl = 2.0 * (b - a);
  xn = xn - dbl * floor((xn - a) / dbl);
  if (xn >= b) xn = 2.0 * b - xn;
  for (int it = 0; it < 8 && (xn <= a || xn >= b); it++) xn = xn >= b ? 2.0 * b - xn : 2.0 * a - xn;
  if (xn <= a || xn >= b) xn = (a + b) / 2.0;
  return xn;
}

__device__ inline double smpTau(const SmpModel& m, int pop, int ovPop, double ovTau) { return pop == ovPop ? ovTau : m.tau[pop]; }
__device__ inline double smpPopEnd(const SmpModel& m, int pop, int ovPop, double ovTau) {
  return m.father[pop] >= 0 ? smpTau(m, m.father[pop], ovPop, ovTau) : kOldAge;
}

// ------------------------------------------------------------------------------------------ a locus held by a warp
// Every sampler kernel gives one WARP to a locus: lane i holds nodes i, i+32, ... (R per lane) in registers, so the
// O(N^2) work of a proposal is spread over the lanes and no per-locus scratch memory is needed.
constexpr int kSmpLociPerCta = kSmpThreads / 32;
constexpr double kSmpInf = 1e300;

template <int R>
struct WarpLocus {
  double age[R];
  int pop[R];      // population of the node; -1 for slots beyond the last node
  int father[R];
};

template <int R>
__device__ inline void wlLoad(WarpLocus<R>& w, const StoreDev& d, const SmpDev& sd, int l, int lane) {
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int x = lane + 32 * r;
    if (x < d.N) {
      const size_t o = (size_t)l * d.N + x;
      w.age[r] = d.age[o];
      w.pop[r] = sd.nodePop[o];
      w.father[r] = d.node[o].father;
    } else {
      w.age[r] = kSmpInf; w.pop[r] = -1; w.father[r] = -1;
    }
  }
}

__device__ inline double warpSumD(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Statistics of the locus (computeGenetreeStats + recalcStats, patch.c:2330-2513): coal[p] = sum over the
// intervals of population p of n(n-1)*length, ncoal[p] = coalescences in p.  For every coalescence x the lanes find,
// in one all-pairs pass over shuffled (age, population) pairs, how many coalescences of the same population precede
// it and the age of the latest one; its interval then contributes n_x(n_x - 1)(age_x - previous).  Sums per
// population are fixed-order shuffle trees (deterministic).  scratch: per-warp shared ints [2*Q].
template <int R>
__device__ inline void wlStats(const SmpModel& m, const WarpLocus<R>& w, int n, int N, int lane, int ovPop, double ovTau,
                               int* scratch, double* coalOut, int* ncoalOut) {
  const int Q = m.Q;
  int* sNcoal = scratch;
  int* sNstart = scratch + Q;
  int ev[R];   // population of the node if it is a coalescence, else -1
#pragma unroll
  for (int r = 0; r < R; r++) ev[r] = (lane + 32 * r >= n && lane + 32 * r < N) ? w.pop[r] : -1;
  for (int p = 0; p < Q; p++) {
    int c = 0;
#pragma unroll
    for (int r = 0; r < R; r++) c += __popc(__ballot_sync(0xffffffffu, ev[r] == p));
    if (lane == 0) sNcoal[p] = c;
  }
  __syncwarp();
  for (int p = lane; p < Q; p += 32) {   // lineages entering p = samples below it minus coalescences strictly below it
    int lin = m.leavesBelow[p];
    for (int q = 0; q < Q; q++)
      if (q != p && ((m.below[p] >> q) & 1ull)) lin -= sNcoal[q];
    sNstart[p] = lin;
  }
  __syncwarp();
  int cnt[R];
  double prev[R];
#pragma unroll
  for (int r = 0; r < R; r++) { cnt[r] = 0; prev[r] = ev[r] >= 0 ? smpTau(m, ev[r], ovPop, ovTau) : 0.0; }
#pragma unroll
  for (int r2 = 0; r2 < R; r2++) {
    for (int src = 0; src < 32; src++) {
      const int y = src + 32 * r2;
      if (y < n) continue;
      if (y >= N) break;
      const double ay = __shfl_sync(0xffffffffu, w.age[r2], src);
      const int py = __shfl_sync(0xffffffffu, ev[r2], src);
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int x = lane + 32 * r;
        if (py == ev[r] && (ay < w.age[r] || (ay == w.age[r] && y < x))) {
          cnt[r]++;
          prev[r] = fmax(prev[r], ay);
        }
      }
    }
  }
  double term[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    term[r] = 0.0;
    if (ev[r] >= 0) {
      const int lin = sNstart[ev[r]] - cnt[r];
      term[r] = (double)(lin * (lin - 1)) * (w.age[r] - prev[r]);
      if (cnt[r] == sNcoal[ev[r]] - 1) {   // last coalescence of its population: the interval up to the population's end
        const int rest = lin - 1;
        term[r] += (double)(rest * (rest - 1)) * (smpPopEnd(m, ev[r], ovPop, ovTau) - w.age[r]);
      }
    }
  }
  for (int p = 0; p < Q; p++) {
    double v = 0.0;
#pragma unroll
    for (int r = 0; r < R; r++) v += ev[r] == p ? term[r] : 0.0;
    v = warpSumD(v);
    if (lane == 0) {
      if (sNcoal[p] == 0) {
        const int lin = sNstart[p];
        v = (double)(lin * (lin - 1)) * (smpPopEnd(m, p, ovPop, ovTau) - smpTau(m, p, ovPop, ovTau));
      }
      coalOut[p] = v;
      ncoalOut[p] = sNcoal[p];
    }
  }
  __syncwarp();
}

// coal statistic of ONE population (a coalescence-time move changes no other): the all-pairs pass of wlStats
// restricted to the coalescences of that population, found with a ballot.  nStart = lineages entering it.
template <int R>
__device__ inline double wlPopStat(const SmpModel& m, const WarpLocus<R>& w, int n, int N, int lane, int pop, int nStart) {
  unsigned mask[R];
  int total = 0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int x = lane + 32 * r;
    mask[r] = __ballot_sync(0xffffffffu, x >= n && x < N && w.pop[r] == pop);
    total += __popc(mask[r]);
  }
  const double tau = m.tau[pop];
  const double end = m.father[pop] >= 0 ? m.tau[m.father[pop]] : kOldAge;
  if (total == 0) return (double)(nStart * (nStart - 1)) * (end - tau);
  int cnt[R];
  double prev[R];
#pragma unroll
  for (int r = 0; r < R; r++) { cnt[r] = 0; prev[r] = tau; }
#pragma unroll
  for (int r2 = 0; r2 < R; r2++) {
    for (unsigned rest = mask[r2]; rest; rest &= rest - 1) {
      const int src = __ffs(rest) - 1;
      const int y = src + 32 * r2;
      const double ay = __shfl_sync(0xffffffffu, w.age[r2], src);
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int x = lane + 32 * r;
        if (ay < w.age[r] || (ay == w.age[r] && y < x)) { cnt[r]++; prev[r] = fmax(prev[r], ay); }
      }
    }
  }
  double v = 0.0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    if ((mask[r] >> lane) & 1u) {
      const int lin = nStart - cnt[r];
      v += (double)(lin * (lin - 1)) * (w.age[r] - prev[r]);
      if (cnt[r] == total - 1) {
        const int restLin = lin - 1;
        v += (double)(restLin * (restLin - 1)) * (end - w.age[r]);
      }
    }
  }
  return warpSumD(v);
}

// change of the genealogy log-density between the stored and the tentative statistics of a locus (lane 0's value)
__device__ inline double smpGenDelta(const SmpModel& m, const double* coalOld, const double* coalNew, const int* ncoalOld,
                                     const int* ncoalNew) {
  double delta = 0.0;
  for (int p = 0; p < m.Q; p++) {
    delta -= (coalNew[p] - coalOld[p]) / m.theta[p];
    if (ncoalNew[p] != ncoalOld[p]) delta += (double)(ncoalNew[p] - ncoalOld[p]) * log(2.0 / m.theta[p]);
  }
  return delta;
}

// the model is read with lane-dependent indices inside dependent loops (population walks): a shared-memory copy per
// CTA instead of global loads
#define SMP_STAGE_MODEL                                                                          \
  __shared__ SmpModel smpModelShared;                                                            \
  for (int i_ = threadIdx.x; i_ < (int)(sizeof(SmpModel) / sizeof(int)); i_ += blockDim.x)       \
    ((int*)&smpModelShared)[i_] = ((const int*)mp)[i_];                                          \
  __syncthreads();

#define SMP_WARP_PROLOGUE                                                       \
  extern __shared__ int smpScratch[];                                           \
  SMP_STAGE_MODEL                                                               \
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;                    \
  const int l = sd.l0 + blockIdx.x * kSmpLociPerCta + wid;                      \
  if (l >= sd.l1) return;                                                       \
  const SmpModel& m = smpModelShared;                                           \
  int* scratch = smpScratch + wid * 2 * kSmpMaxPops;                            \
  const TreeView t = deviceView(d, l);                                          \
  const int n = d.n, N = d.N;                                                   \
  (void)scratch; (void)n; (void)N; (void)lane; (void)m;

__device__ inline SmpProposal smpNoProposal() {
  SmpProposal pr;
  pr.genDelta = 0.0; pr.aux = 0.0; pr.pop = 0; pr.node = -1; pr.valid = 0; pr.ntj0 = 0; pr.ntj1 = 0;
  return pr;
}

__device__ inline void smpResolve(const StoreDev& d, const SmpDev& sd, const SmpModel& m, const TreeView& t, int l, int lane, int N,
                                  int kind, unsigned long long seed, unsigned long long step);

// ------------------------------------------------------------------------------------------ coalescence-time move
template <int R>
__device__ inline void smpAgeProposeBody(const StoreDev& d, const SmpDev& sd, const SmpModel& m, const TreeView& t, int l, int lane,
                                         int n, int N, int inode, double finetune, unsigned long long seed, unsigned long long step) {
  SmpProposal pr = smpNoProposal();
  pr.node = inode;
  const int root = *t.root;
  if (root < n) { if (lane == 0) sd.prop[l] = pr; return; }
  WarpLocus<R> w;
  wlLoad(w, d, sd, l, lane);
  double tnew = 0.0;
  int valid = 0;
  if (lane == 0) {
    const uint8_t* np = sd.nodePop + (size_t)l * N;
    const int pop = np[inode];
    const double told = t.age[inode];
    const NodeRec rec = t.node[inode];
    const double lo = fmax(m.tau[pop], fmax(t.age[rec.left], t.age[rec.right]));
    double hi = m.father[pop] >= 0 ? m.tau[m.father[pop]] : kOldAge;
    if (inode != root) hi = fmin(hi, t.age[rec.father]);
    SmpRng rng(seed, (unsigned long long)l, step);
    tnew = smpReflect(told + finetune * rng.normal2(), lo, hi);
    valid = fabs(tnew - told) >= 1e-15;   // GPhoCS.c:2354-2358
    if (valid) adjustAge(t, inode, tnew);
  }
  valid = __shfl_sync(0xffffffffu, valid, 0);
  tnew = __shfl_sync(0xffffffffu, tnew, 0);
  if (valid) {
    int pop = 0, nStart = 0;
    if (lane == 0) {
      pop = sd.nodePop[(size_t)l * N + inode];
      const int* nc = sd.ncoal + (size_t)l * m.Q;
      nStart = m.leavesBelow[pop];
      for (int q = 0; q < m.Q; q++)
        if (q != pop && ((m.below[pop] >> q) & 1ull)) nStart -= nc[q];
    }
    pop = __shfl_sync(0xffffffffu, pop, 0);
    nStart = __shfl_sync(0xffffffffu, nStart, 0);
#pragma unroll
    for (int r = 0; r < R; r++)
      if (lane + 32 * r == inode) w.age[r] = tnew;
    const double coalNew = wlPopStat<R>(m, w, n, N, lane, pop, nStart);
    if (lane == 0) {
      pr.genDelta = -(coalNew - sd.coal[(size_t)l * m.Q + pop]) / m.theta[pop];
      pr.aux = coalNew;
      pr.pop = pop;
      pr.valid = 1;
    }
  }
  if (lane == 0) sd.prop[l] = pr;
}
template <int R>
__global__ void __launch_bounds__(kSmpThreads)
k_smp_age_propose(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int inode, double finetune, unsigned long long seed,
                  unsigned long long step, int pendKind, unsigned long long pendStep) {
  SMP_WARP_PROLOGUE
  if (pendKind >= 0) smpResolve(d, sd, m, t, l, lane, N, pendKind, seed, pendStep);   // the previous proposal of this locus
  smpAgeProposeBody<R>(d, sd, m, t, l, lane, n, N, inode, finetune, seed, step);
}

// ------------------------------------------------------------------------------------------ subtree prune and regraft
// The pruned lineage is re-attached by the coalescent conditional on the rest of the genealogy: while it shares
// a population with another lineage it coalesces with it at rate 2/theta of that population.  These are
// independent competing clocks, so every lane draws the first ring of its own lineages' clocks (integrating the
// piecewise-constant rate along the populations the pruned lineage passes through) and the earliest ring over
// the warp gives the new coalescence time, its population and the target branch — the same law as walking the
// intervals of the event chain (traceLineage, patch.c:886-1331, without migration).  The proposal is the
// conditional prior, so the acceptance ratio is the data-likelihood ratio alone (GPhoCS.c:2702-2706).
template <int R>
__device__ inline void smpSprProposeBody(const StoreDev& d, const SmpDev& sd, const SmpModel& m, const TreeView& t, int l, int lane,
                                         int n, int N, int node, unsigned long long seed, unsigned long long step) {
  SmpProposal pr = smpNoProposal();
  const int root = *t.root;
  if (root < n || node == root) { if (lane == 0) sd.prop[l] = pr; return; }
  WarpLocus<R> w;
  wlLoad(w, d, sd, l, lane);
  const int F = t.node[node].father;
  const NodeRec recF = t.node[F];
  const int S = recF.left + recF.right - node;
  const int G = recF.father;
  const double t0 = t.age[node];
  const int pop0 = sd.nodePop[(size_t)l * N + node];
  const double ageG = G >= 0 ? t.age[G] : kSmpInf;
  double bestT = kSmpInf;
  int bestX = -1, bestPop = -1;
  const SmpRng rng(seed, (unsigned long long)l, step);
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int x = lane + 32 * r;
    if (x >= N || x == node || x == F) continue;
    const double endx = x == S ? ageG : (w.father[r] >= 0 ? t.age[w.father[r]] : kSmpInf);
    int q = w.pop[r];   // first population in which the two lineages can meet: their common ancestor
    while (!((m.below[q] >> pop0) & 1ull)) q = m.father[q];
    double sNow = fmax(fmax(t0, w.age[r]), m.tau[q]);
    if (sNow >= endx) continue;
    while (m.father[q] >= 0 && m.tau[m.father[q]] <= sNow) q = m.father[q];
    double need = rng.exponentialAt((unsigned long long)x);
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
  for (int off = 16; off > 0; off >>= 1) {   // earliest ring over the warp (ties: lower node id)
    const double oT = __shfl_xor_sync(0xffffffffu, bestT, off);
    const int oX = __shfl_xor_sync(0xffffffffu, bestX, off);
    const int oP = __shfl_xor_sync(0xffffffffu, bestPop, off);
    if (oT < bestT || (oT == bestT && oX >= 0 && (bestX < 0 || oX < bestX))) { bestT = oT; bestX = oX; bestPop = oP; }
  }
  if (bestX >= 0) {
    if (lane == 0) {
      uint8_t* np = sd.nodePop + (size_t)l * N;
      pr.pop = np[F];
      pr.node = F;
      spr(t, node, bestX, bestT);
      np[F] = (uint8_t)bestPop;
    }
    pr.valid = 1;   // the statistics of accepted loci are refreshed once, after the sweep (k_smp_init_stats)
  }
  if (lane == 0) sd.prop[l] = pr;
}
template <int R>
__global__ void __launch_bounds__(kSmpThreads)
k_smp_spr_propose(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int node, unsigned long long seed,
                  unsigned long long step, int pendKind, unsigned long long pendStep) {
  SMP_WARP_PROLOGUE
  if (pendKind >= 0) smpResolve(d, sd, m, t, l, lane, N, pendKind, seed, pendStep);
  smpSprProposeBody<R>(d, sd, m, t, l, lane, n, N, node, seed, step);
}

// ------------------------------------------------------------------------------------------ per-locus accept / reject
// kind 0: coalescence-time move (likelihood ratio of data and genealogy); kind 1: SPR (data likelihood ratio).
// A coalescence-time move changes one population's statistic (carried in the proposal record); after an SPR
// sweep the statistics of all loci are recomputed in one launch.
__device__ inline void smpResolve(const StoreDev& d, const SmpDev& sd, const SmpModel& m, const TreeView& t, int l, int lane, int N,
                                  int kind, unsigned long long seed, unsigned long long step) {
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
      if (lane == 0) {
        commitLocus(t);
        if (kind == 0) sd.coal[(size_t)l * m.Q + pr.pop] = pr.aux;
      }
    } else {
      for (int x = lane; x < N; x += 32) revertNode(t, x);
      if (lane == 0) {
        revertLocus(t);
        if (kind == 1) sd.nodePop[(size_t)l * N + pr.node] = (uint8_t)pr.pop;
      }
    }
  } else if (kind == 0) {
    ok = 1;   // an unchanged age counts as accepted (GPhoCS.c:2354-2358)
  }
  if (lane == 0 && ok) atomicAdd(sd.accepted + kind, 1ull);
  __syncwarp();   // the lanes that follow read what other lanes have just committed or reverted
}
__global__ void __launch_bounds__(kSmpThreads)
k_smp_accept(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int kind, unsigned long long seed, unsigned long long step) {
  SMP_WARP_PROLOGUE
  smpResolve(d, sd, m, t, l, lane, N, kind, seed, step);
}

// ------------------------------------------------------------------------------------------ split-time move
// Rubber band (patch.c:596-801) around population A whose split time moves tauOld -> tauNew: coalescences of A
// in (tauOld, ub) are rescaled towards ub by f1, coalescences of its two sons in (lb, tauOld) towards lb by f0; in
// the root population everything above tauOld is rescaled from lb by f0 (GPhoCS.c:3749-3785).
template <int R>
__global__ void __launch_bounds__(kSmpThreads)
k_smp_tau_propose(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int A, double tauOld, double tauNew, double lb, double ub,
                  double f0, double f1) {
  SMP_WARP_PROLOGUE
  SmpProposal pr = smpNoProposal();
  pr.pop = A;
  if (*t.root < n) { if (lane == 0) sd.prop[l] = pr; return; }
  WarpLocus<R> w;
  wlLoad(w, d, sd, l, lane);
  const bool isRoot = A == m.rootPop;
  // A current population: its SAMPLE AGE moves (UpdateSampleAge, GPhoCS.c:4006-4590) — the leaves of A take the new
  // age, the coalescences of A above it are rubber-banded towards the population's end; there are no sons.
  const int s0 = A >= m.C ? m.son0[A] : -1, s1 = A >= m.C ? m.son1[A] : -1;
  int n0 = 0, n1 = 0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int x = lane + 32 * r;
    int which = 0;   // 1: lower band, 2: upper band, 3: sample of A
    if (x >= n && x < N) {
      const int q = w.pop[r];
      const double a = w.age[r];
      if (q == A) { if (isRoot || (a > tauOld && a < ub)) which = 2; }
      else if ((q == s0 || q == s1) && a > lb && a < tauOld) which = 1;
    } else if (x < n && A < m.C && w.pop[r] == A) {
      which = 3;
    }
    if (which) {
      const double a = w.age[r];
      const double an = which == 3 ? tauNew : (which == 1 || isRoot ? lb + (a - lb) * f0 : ub + (a - ub) * f1);
      adjustAge(t, x, an);
      w.age[r] = an;
    }
    n0 += __popc(__ballot_sync(0xffffffffu, which == 1));
    n1 += __popc(__ballot_sync(0xffffffffu, which == 2));
  }
  double* cT = sd.coalT + (size_t)l * m.Q;
  int* nT = sd.ncoalT + (size_t)l * m.Q;
  wlStats<R>(m, w, n, N, lane, A, tauNew, scratch, cT, nT);
  if (lane == 0) {
    pr.genDelta = smpGenDelta(m, sd.coal + (size_t)l * m.Q, cT, sd.ncoal + (size_t)l * m.Q, nT);
    pr.ntj0 = n0;
    pr.ntj1 = n1;
    pr.valid = 1;
    sd.prop[l] = pr;
  }
}

// joint rescaling of every node age by c (scaleAllNodeAges, LocusDataLikelihood.c:895-917, without its evaluation)
__global__ void __launch_bounds__(kSmpThreads) k_smp_scale_propose(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, double c) {
  SMP_WARP_PROLOGUE
  SmpProposal pr = smpNoProposal();
  if (*t.root >= n) {
    for (int x = lane; x < N; x += 32) adjustAge(t, x, c * t.age[x]);
    pr.valid = 1;
  }
  if (lane == 0) sd.prop[l] = pr;
}

// block partial sums, V = 5 + 2Q + 2B values:
//   [0] data delta (mode 0) or sum of data lnL (mode 1), [1] genealogy delta, [2] ntj0, [3] ntj1, [4] loci in conflict,
//   [5..5+Q) coal totals, [5+Q..5+2Q) ncoal totals, [5+2Q..+B) mig totals, [..+B) nmig totals   (mode 1 only)
__global__ void __launch_bounds__(kSmpThreads) k_smp_reduce(StoreDev d, SmpDev sd, int mode, int B) {
  __shared__ double sh[kSmpThreads];
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int Q = sd.Q;
  const int V = mode == 0 ? 5 : 5 + 2 * Q + 2 * B;
  for (int v = 0; v < V; v++) {
    double x = 0.0;
    if (l < d.L) {
      if (mode == 0) {
        const SmpProposal pr = sd.prop[l];
        if (v == 0) x = pr.valid ? d.lnL[l] - d.savedLnL[l] : 0.0;
        else if (v == 1) x = pr.genDelta;
        else if (v == 2) x = pr.ntj0;
        else if (v == 3) x = pr.ntj1;
        else x = pr.valid && pr.node == -2 ? 1.0 : 0.0;
      } else {
        if (v == 0) x = d.lnL[l];
        else if (v == 1) x = (d.rate[l] - 1.0) * (d.rate[l] - 1.0);
        else if (v == 2) x = 1.0;   // loci (of all ranks after the all-reduce)
        else if (v >= 5 && v < 5 + Q) x = sd.coal[(size_t)l * Q + (v - 5)];
        else if (v >= 5 + Q && v < 5 + 2 * Q) x = sd.ncoal[(size_t)l * Q + (v - 5 - Q)];
        else if (v >= 5 + 2 * Q && v < 5 + 2 * Q + B) x = sd.mig[(size_t)l * B + (v - 5 - 2 * Q)];
        else if (v >= 5 + 2 * Q + B) x = sd.nmig[(size_t)l * B + (v - 5 - 2 * Q - B)];
      }
    }
    sh[threadIdx.x] = x;
    __syncthreads();
    for (int off = kSmpThreads / 2; off > 0; off >>= 1) {
      if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
      __syncthreads();
    }
    if (threadIdx.x == 0) sd.partial[(size_t)blockIdx.x * kSmpPartials + v] = sh[0];
    __syncthreads();
  }
}

// total[v] = sum over blocks of partial[block][v] in block order (v < nv; the rest 0): one thread per value
__global__ void __launch_bounds__(128) k_smp_reduce_final(const double* __restrict__ partial, int blocks, int nv, int V,
                                                          double* __restrict__ total) {
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    double acc = 0.0;
    if (v < nv)
      for (int b = 0; b < blocks; b++) acc += partial[(size_t)b * kSmpPartials + v];
    total[v] = acc;
  }
}

// one global accept / reject for every locus; how: 0 tau move (statistics <- tentative), 1 rescaling (statistics *= c)
__global__ void __launch_bounds__(kSmpThreads) k_smp_global_resolve(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int accept,
                                                                    int how, double c) {
  SMP_WARP_PROLOGUE
  if (!sd.prop[l].valid) return;
  if (accept) {
    for (int x = lane; x < N; x += 32) commitNode(t, x);
    for (int p = lane; p < sd.Q; p += 32) {
      double* c0 = sd.coal + (size_t)l * sd.Q + p;
      *c0 = how == 0 ? sd.coalT[(size_t)l * sd.Q + p] : *c0 * c;
    }
    if (lane == 0) commitLocus(t);
  } else {
    for (int x = lane; x < N; x += 32) revertNode(t, x);
    if (lane == 0) revertLocus(t);
  }
}

// statistics of every locus from scratch (initialisation and consistency checks)
template <int R>
__global__ void __launch_bounds__(kSmpThreads) k_smp_init_stats(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int toTentative) {
  SMP_WARP_PROLOGUE
  if (*t.root < n) return;
  WarpLocus<R> w;
  wlLoad(w, d, sd, l, lane);
  wlStats<R>(m, w, n, N, lane, -1, 0.0, scratch, (toTentative ? sd.coalT : sd.coal) + (size_t)l * m.Q,
             (toTentative ? sd.ncoalT : sd.ncoal) + (size_t)l * m.Q);
}

// ------------------------------------------------------------------------------------------ locus-rate moves
// UpdateLocusRate (GPhoCS.c:4598-4675) keeps the mean rate at 1 by moving rate between a locus and a reference locus,
// one locus after the other.  Here the same move — rate shifted between TWO loci, their sum kept, Dirichlet(alpha)
// prior ratio, both data likelihoods recomputed from scratch — is made on disjoint pairs (g, g + offset) for all
// pairs at once; the offset changes every iteration so rate can travel between any two loci.
// Pairs are formed over the GLOBAL locus index (all ranks): the two loci of a pair may live on different GPUs.  Every
// rank sees the gathered rates (before the move) and, after the evaluation, the gathered log-likelihood changes of all
// loci, and computes proposal and decision of a pair from the same numbers with the same random stream (keyed by the
// pair's lower global index): both owners act alike without talking to each other.
// Loci are first rotated by `shift` (so that every pair of loci can meet), then (g', g' + offset) pair up inside
// blocks of 2 * offset.  Returns the partner (-1: unpaired this round); *proposer = this locus is the lower of its pair.
__device__ inline long long smpRatePartner(long long g, long long Lg, long long offset, long long shift, bool* proposer) {
  const long long gp = (g + shift) % Lg;
  const bool lower = ((gp / offset) & 1) == 0;
  *proposer = lower;
  const long long pp = lower ? gp + offset : gp - offset;
  if (pp >= Lg) return -1;
  return (pp - shift + Lg) % Lg;
}
// gather[g] = rate of global locus g, negative if the locus has no genealogy (never moved); own segment only — the
// all-reduce that follows fills in the other ranks' (their segments are zero here)
__global__ void __launch_bounds__(kSmpThreads) k_smp_rate_gather(StoreDev d, double* __restrict__ gather, long long off) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < d.L) gather[off + l] = d.root[l] >= d.n ? d.rate[l] : -d.rate[l];
}
// the pair's proposal, from the gathered rates: new rate of the lower locus; valid = both loci can move
__device__ inline double smpRateProposal(const double* __restrict__ gather, long long lo, long long hi, double finetune,
                                         unsigned long long seed, unsigned long long step, bool* valid) {
  const double r0 = gather[lo], r1 = gather[hi];
  *valid = r0 > 0.0 && r1 > 0.0;
  if (!*valid) return 0.0;
  SmpRng rng(seed, (unsigned long long)lo, step);
  return smpReflect(r0 + finetune * rng.normal2(), 0.0, r0 + r1);
}
__global__ void __launch_bounds__(kSmpThreads)
k_smp_rate_propose(StoreDev d, const double* __restrict__ gather, long long off, long long Lg, long long offset, long long shift,
                   double finetune, unsigned long long seed, unsigned long long step) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= d.L) return;
  bool proposer, valid;
  const long long g = off + l, partner = smpRatePartner(g, Lg, offset, shift, &proposer);
  if (partner < 0) return;
  const long long lo = proposer ? g : partner, hi = proposer ? partner : g;
  const double rn = smpRateProposal(gather, lo, hi, finetune, seed, step, &valid);
  if (valid) d.rate[l] = proposer ? rn : gather[lo] + gather[hi] - rn;
}
// after a full evaluation (useOld = 0) of every locus: gather[g] = change of the locus' log-likelihood (own segment)
__global__ void __launch_bounds__(kSmpThreads) k_smp_rate_delta(StoreDev d, double* __restrict__ gather, long long off) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < d.L) gather[off + l] = d.lnL[l] - d.savedLnL[l];
}
// both owners of a pair reach the same decision from the same numbers, then commit or put the old rate back
__global__ void __launch_bounds__(kSmpThreads)
k_smp_rate_resolve(StoreDev d, SmpDev sd, const double* __restrict__ rates, const double* __restrict__ delta, long long off, long long Lg,
                   long long offset, long long shift, double finetune, double alpha, unsigned long long seed, unsigned long long step) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= d.L) return;
  bool proposer, valid = false;
  const long long g = off + l, partner = smpRatePartner(g, Lg, offset, shift, &proposer);
  int ok = 0;
  if (partner >= 0) {
    const long long lo = proposer ? g : partner, hi = proposer ? partner : g;
    const double rn = smpRateProposal(rates, lo, hi, finetune, seed, step, &valid);
    if (valid) {
      const double r0 = rates[lo], r1 = rates[hi];
      double lnacc = delta[lo] + delta[hi];
      lnacc += (alpha - 1.0) * log((rn * (r0 + r1 - rn)) / (r0 * r1));
      ok = lnacc >= 0.0;
      if (!ok) {
        SmpRng rng(seed, (unsigned long long)lo, step + 1ull);
        ok = rng.uniform() < exp(lnacc);
      }
    }
  }
  const TreeView t = deviceView(d, l);
  if (ok) {
    commit(t);
    if (proposer) atomicAdd(sd.accepted + 2, 1ull);
  } else {
    revert(t);   // also for unpaired loci: the full evaluation flipped their buffers
    if (valid) d.rate[l] = rates[g];
  }
}

// consistency of the population assignment: every coalescence lies inside its population's time span and above
// both children, in a population that is the children's or an ancestor of it; returns violations per locus
__global__ void __launch_bounds__(kSmpThreads) k_smp_check(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int* __restrict__ bad) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= d.L) return;
  const SmpModel& m = *mp;
  const TreeView t = deviceView(d, l);
  int v = 0;
  if (*t.root >= d.n) {
    const uint8_t* np = sd.nodePop + (size_t)l * d.N;
    for (int x = d.n; x < d.N; x++) {
      const int q = np[x];
      const double a = t.age[x];
      const double end = m.father[q] >= 0 ? m.tau[m.father[q]] : kOldAge;
      if (!(a >= m.tau[q] && a <= end)) v++;
      const NodeRec r = t.node[x];
      if (!(t.age[r.left] <= a && t.age[r.right] <= a)) v++;
      // without migration a coalescence happens in its children's population or an ancestor of it
      if (m.B == 0 && (!((m.below[q] >> np[r.left]) & 1ull) || !((m.below[q] >> np[r.right]) & 1ull))) v++;
      if (r.father >= 0 ? t.node[r.father].left != x && t.node[r.father].right != x : *t.root != x) v++;
      if (t.node[x].flags & (F_RECALC | F_SAVED)) v++;
    }
  }
  bad[l] = v;
}

}  // namespace gphocs
