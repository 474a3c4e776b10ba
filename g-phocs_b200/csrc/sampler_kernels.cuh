// sampler_kernels.cuh — device-resident MCMC update steps (proposals on the device).
//
// What the reference does on the host, one locus at a time, with a device round trip per proposal
//   UpdateGB_InternalNode  GPhoCS.c:2287-2424   coalescence-time moves
//   UpdateGB_MigSPR        GPhoCS.c:2598-2760   subtree prune + re-coalescence drawn from the coalescent prior
//   UpdateTheta            GPhoCS.c:3035-3103   multiplicative theta moves from the total statistics
//   UpdateTau              GPhoCS.c:3224-3990   split-time moves with the rubber band (patch.c:596-801)
//   mixing                 GPhoCS.c:4688-4900   joint rescaling of all times and thetas
// is done here for ALL loci per launch: one thread per locus proposes (its own counter-based Philox stream) and
// edits the device-resident genealogy through the same edit protocol the host API uses (tree_ops.cuh); k_eval
// (clv_kernels.cuh) evaluates every locus incrementally; one thread per locus accepts or rejects.  No host round
// trip inside a sweep.  Scope of this version: population trees WITHOUT migration bands, samples of age 0,
// constant locus rates — the model of BASELINE.json configs[1].  The target density is the reference's:
//   P(X|G) * prod_pops (2/theta)^ncoal exp(-coal_stats/theta) * Gamma priors on theta and tau
// (gtreeLnLikelihood patch.c:2702-2723; priors as used in GPhoCS.c:3063-3065, 3444-3447, 4741-4760).
// The chains are not bit-identical to the reference's (different random streams): parity is statistical
// (tests/test_gpu_sampler.py: prior recovery with uninformative data, posterior means against the reference chain).
#pragma once
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdint.h>

#include "clv_kernels.cuh"

namespace gphocs {

constexpr int kSmpMaxPops = 39;     // 2*NSPECIES-1 (patch.h:19)
constexpr double kOldAge = 999.0;   // OLDAGE (patch.h:22): end of the root population

struct SmpModel {
  int Q, C, rootPop, n;
  int father[kSmpMaxPops], son0[kSmpMaxPops], son1[kSmpMaxPops], samplesPerPop[kSmpMaxPops];
  int postOrder[kSmpMaxPops];
  int leavesBelow[kSmpMaxPops];       // haploid samples in the current populations under each population
  unsigned long long below[kSmpMaxPops];  // bit q: population q is this population or lies below it
  double theta[kSmpMaxPops], tau[kSmpMaxPops];
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
  uint8_t* nodePop;    // [L][N] population of every genealogy node (nodePops, patch.h:123)
  double* coal;        // [L][Q] coal_stats per locus
  double* coalT;       // [L][Q] tentative statistics of a pending global proposal
  int* ncoal;          // [L][Q] num_coals per locus
  SmpProposal* prop;   // [L]
  unsigned long long* accepted;  // [8] acceptance counters per move kind
  double* partial;     // [blocks][kSmpPartials] block partial sums
};
constexpr int kSmpPartials = 4 + 2 * kSmpMaxPops;
constexpr int kSmpThreads = 128;

struct SmpRng {
  curandStatePhilox4_32_10_t st;
  __device__ SmpRng(unsigned long long seed, unsigned long long locus, unsigned long long step) {
    curand_init(seed, locus, step * 16ull, &st);   // 16 values reserved per (locus, step)
  }
  __device__ double uniform() { return curand_uniform_double(&st); }   // (0, 1]
  __device__ double normal() { return curand_normal_double(&st); }
  // rnd2normal8 (utils.c:482-488): mixture of N(-m, s^2) and N(m, s^2) with m^2 + s^2 = 1, m^2/s^2 = 8
  __device__ double normal2() {
    const double z = 0.94280904158206336 + normal() * (1.0 / 3.0);
    return uniform() < 0.5 ? z : -z;
  }
  __device__ double exponential() { return -log(uniform()); }
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

// coal statistic of one population: sum over the intervals between its coalescence events of n(n-1)*length
// (recalcStats, patch.c:2403-2413), events visited in age order by repeated selection (no scratch memory).
// nStart = lineages entering the population.  Returns the number of coalescences through *numCoals.
__device__ inline double smpPopCoalStat(const SmpModel& m, const double* age, const uint8_t* nodePop, int pop, int nStart,
                                        int ovPop, double ovTau, int* numCoals) {
  const int n = m.n, N = 2 * n - 1;
  double t = smpTau(m, pop, ovPop, ovTau);
  const double tEnd = smpPopEnd(m, pop, ovPop, ovTau);
  double stat = 0.0;
  int lin = nStart, prev = -1, count = 0;
  double prevAge = -1.0;
  for (;;) {
    // next coalescence of this population after (prevAge, prev) in (age, id) order
    int best = -1;
    double bestAge = 0.0;
    for (int x = n; x < N; x++) {
      if (nodePop[x] != pop) continue;
      const double a = age[x];
      if (a < prevAge || (a == prevAge && x <= prev)) continue;
      if (best < 0 || a < bestAge || (a == bestAge && x < best)) { best = x; bestAge = a; }
    }
    if (best < 0) break;
    stat += (double)(lin * (lin - 1)) * (bestAge - t);
    t = bestAge;
    lin--;
    count++;
    prev = best;
    prevAge = bestAge;
  }
  stat += (double)(lin * (lin - 1)) * (tEnd - t);
  if (numCoals) *numCoals = count;
  return stat;
}

// lineages entering a population, from the coalescence counts of the populations below it
__device__ inline int smpLineagesEntering(const SmpModel& m, const int* ncoal, int pop) {
  int lin = m.leavesBelow[pop];
  for (int q = 0; q < m.Q; q++)
    if (q != pop && ((m.below[pop] >> q) & 1ull)) lin -= ncoal[q];
  return lin;
}

// all statistics of a locus from its genealogy (computeGenetreeStats, patch.c:2330-2354)
__device__ inline void smpLocusStats(const SmpModel& m, const double* age, const uint8_t* nodePop, int ovPop, double ovTau,
                                     double* coal, int* ncoal) {
  int nEnd[kSmpMaxPops];
  for (int i = 0; i < m.Q; i++) {
    const int p = m.postOrder[i];
    const int nStart = p < m.C ? m.samplesPerPop[p] : nEnd[m.son0[p]] + nEnd[m.son1[p]];
    int nc = 0;
    coal[p] = smpPopCoalStat(m, age, nodePop, p, nStart, ovPop, ovTau, &nc);
    ncoal[p] = nc;
    nEnd[p] = nStart - nc;
  }
}

// population in which a lineage that started in `pop` lives at time s
__device__ inline int smpPopAt(const SmpModel& m, int pop, double s) {
  while (m.father[pop] >= 0 && m.tau[m.father[pop]] <= s) pop = m.father[pop];
  return pop;
}

// ------------------------------------------------------------------------------------------ coalescence-time move
__global__ void __launch_bounds__(kSmpThreads)
k_smp_age_propose(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int inode, double finetune, unsigned long long seed,
                  unsigned long long step) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= d.L) return;
  const SmpModel& m = *mp;
  SmpProposal pr;
  pr.genDelta = 0.0; pr.aux = 0.0; pr.pop = 0; pr.node = inode; pr.valid = 0; pr.ntj0 = 0; pr.ntj1 = 0;
  const TreeView t = deviceView(d, l);
  const int root = *t.root;
  if (root >= d.n && t.numPatterns >= 0) {
    const uint8_t* np = sd.nodePop + (size_t)l * d.N;
    const int pop = np[inode];
    const double told = t.age[inode];
    const NodeRec rec = t.node[inode];
    double lo = fmax(m.tau[pop], fmax(t.age[rec.left], t.age[rec.right]));
    double hi = m.father[pop] >= 0 ? m.tau[m.father[pop]] : kOldAge;
    if (inode != root) hi = fmin(hi, t.age[rec.father]);
    SmpRng rng(seed, (unsigned long long)l, step);
    const double tnew = smpReflect(told + finetune * rng.normal2(), lo, hi);
    if (fabs(tnew - told) >= 1e-15) {   // GPhoCS.c:2354-2358
      const int* nc = sd.ncoal + (size_t)l * m.Q;
      const int nStart = smpLineagesEntering(m, nc, pop);
      adjustAge(t, inode, tnew);
      const double coalNew = smpPopCoalStat(m, t.age, np, pop, nStart, -1, 0.0, nullptr);
      pr.genDelta = -(coalNew - sd.coal[(size_t)l * m.Q + pop]) / m.theta[pop];
      pr.aux = coalNew;
      pr.pop = pop;
      pr.valid = 1;
    }
  }
  sd.prop[l] = pr;
}

// ------------------------------------------------------------------------------------------ subtree prune and regraft
// The pruned lineage is re-attached by simulating the coalescent conditional on the rest of the genealogy: in
// population p with k other lineages it coalesces at rate 2k/theta_p, moves to the parent population at its
// end, and picks its target uniformly among the k lineages.  The proposal is the conditional prior, so the
// acceptance ratio is the data-likelihood ratio alone (GPhoCS.c:2702-2706).
__global__ void __launch_bounds__(kSmpThreads)
k_smp_spr_propose(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int node, unsigned long long seed,
                  unsigned long long step) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= d.L) return;
  const SmpModel& m = *mp;
  SmpProposal pr;
  pr.genDelta = 0.0; pr.aux = 0.0; pr.pop = 0; pr.node = -1; pr.valid = 0; pr.ntj0 = 0; pr.ntj1 = 0;
  const TreeView t = deviceView(d, l);
  const int root = *t.root, n = d.n, N = d.N;
  if (root >= n && node != root) {
    uint8_t* np = sd.nodePop + (size_t)l * N;
    const int F = t.node[node].father;
    const int S = t.node[F].left + t.node[F].right - node;
    const int G = t.node[F].father;
    SmpRng rng(seed, (unsigned long long)l, step);
    double now = t.age[node];
    int pop = smpPopAt(m, np[node], now);
    double need = rng.exponential();
    int target = -1;
    for (int it = 0; it < 4 * N + 4 * kSmpMaxPops && target < 0; it++) {
      // lineages of the pruned genealogy present in `pop` just after `now`, and the next time anything changes
      const double popEnd = m.father[pop] >= 0 ? m.tau[m.father[pop]] : kOldAge * 1e6;
      double next = popEnd;
      int k = 0;
      for (int x = 0; x < N; x++) {
        if (x == node || x == F) continue;
        const double a = t.age[x];
        if (a > now) { next = fmin(next, a); continue; }
        const int par = x == S ? G : t.node[x].father;
        if (par >= 0 && t.age[par] <= now) continue;
        if (smpPopAt(m, np[x], now) == pop) k++;
      }
      const double rate = 2.0 * k / m.theta[pop];
      const double span = next - now;
      if (rate * span >= need) {
        now += need / rate;
        int pick = min(k - 1, (int)(rng.uniform() * k));
        if (pick < 0) pick = 0;
        for (int x = 0; x < N && target < 0; x++) {   // same enumeration as above, at the same reference time
          if (x == node || x == F) continue;
          const double a = t.age[x];
          if (a > now) continue;
          const int par = x == S ? G : t.node[x].father;
          if (par >= 0 && t.age[par] <= now) continue;
          if (smpPopAt(m, np[x], now) != pop) continue;
          if (pick-- == 0) target = x;
        }
        break;
      }
      need -= rate * span;
      now = next;
      if (next >= popEnd && m.father[pop] >= 0) pop = m.father[pop];
    }
    if (target >= 0) {
      pr.pop = np[F];
      pr.node = F;
      spr(t, node, target, now);
      np[F] = (uint8_t)pop;
      pr.valid = 1;
    }
  }
  sd.prop[l] = pr;
}

// ------------------------------------------------------------------------------------------ per-locus accept / reject
// kind 0: coalescence-time move (likelihood ratio of data and genealogy); kind 1: SPR (data likelihood ratio)
__global__ void __launch_bounds__(kSmpThreads)
k_smp_accept(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int kind, unsigned long long seed, unsigned long long step) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long acc = 0;
  if (l < d.L) {
    const SmpModel& m = *mp;
    const SmpProposal pr = sd.prop[l];
    if (pr.valid) {
      const TreeView t = deviceView(d, l);
      const double lnacc = (*t.lnL - *t.savedLnL) + pr.genDelta;
      SmpRng rng(seed, (unsigned long long)l, step);
      bool ok = lnacc >= 0.0;
      if (!ok) ok = rng.uniform() < exp(lnacc);
      if (ok) {
        commit(t);
        if (kind == 0) {
          sd.coal[(size_t)l * m.Q + pr.pop] = pr.aux;
        } else {
          smpLocusStats(m, t.age, sd.nodePop + (size_t)l * d.N, -1, 0.0, sd.coal + (size_t)l * m.Q, sd.ncoal + (size_t)l * m.Q);
        }
        acc = 1;
      } else {
        revert(t);
        if (kind == 1) sd.nodePop[(size_t)l * d.N + pr.node] = (uint8_t)pr.pop;
      }
    } else if (kind == 0) {
      acc = 1;   // an unchanged age counts as accepted (GPhoCS.c:2354-2358)
    }
  }
  acc = __reduce_add_sync(0xffffffffu, (unsigned)acc);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(sd.accepted + kind, acc);
}

// ------------------------------------------------------------------------------------------ split-time move
// Rubber band (patch.c:596-801) around population A whose split time moves tauOld -> tauNew: coalescences of A
// in (tauOld, ub) are rescaled towards ub by f1, coalescences of its two sons in (lb, tauOld) towards lb by f0; in
// the root population everything above tauOld is rescaled from lb by f0 (GPhoCS.c:3749-3785).
__global__ void __launch_bounds__(kSmpThreads)
k_smp_tau_propose(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int A, double tauOld, double tauNew, double lb, double ub,
                  double f0, double f1) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= d.L) return;
  const SmpModel& m = *mp;
  SmpProposal pr;
  pr.genDelta = 0.0; pr.aux = 0.0; pr.pop = A; pr.node = -1; pr.valid = 0; pr.ntj0 = 0; pr.ntj1 = 0;
  const TreeView t = deviceView(d, l);
  if (*t.root >= d.n) {
    const uint8_t* np = sd.nodePop + (size_t)l * d.N;
    const bool isRoot = A == m.rootPop;
    const int s0 = m.son0[A], s1 = m.son1[A];
    for (int x = d.n; x < d.N; x++) {
      const int q = np[x];
      const double a = t.age[x];
      if (q == A) {
        if (isRoot) { adjustAge(t, x, lb + (a - lb) * f0); pr.ntj1++; }
        else if (a > tauOld && a < ub) { adjustAge(t, x, ub + (a - ub) * f1); pr.ntj1++; }
      } else if ((q == s0 || q == s1) && a > lb && a < tauOld) {
        adjustAge(t, x, lb + (a - lb) * f0);
        pr.ntj0++;
      }
    }
    int nc[kSmpMaxPops];
    double* cT = sd.coalT + (size_t)l * m.Q;
    smpLocusStats(m, t.age, np, A, tauNew, cT, nc);
    const double* c0 = sd.coal + (size_t)l * m.Q;
    double delta = 0.0;
    for (int p = 0; p < m.Q; p++) delta -= (cT[p] - c0[p]) / m.theta[p];
    pr.genDelta = delta;
    pr.valid = 1;
  }
  sd.prop[l] = pr;
}

// joint rescaling of every node age by c (scaleAllNodeAges, LocusDataLikelihood.c:895-917, without its evaluation)
__global__ void __launch_bounds__(kSmpThreads) k_smp_scale_propose(StoreDev d, SmpDev sd, double c) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= d.L) return;
  const TreeView t = deviceView(d, l);
  SmpProposal pr;
  pr.genDelta = 0.0; pr.aux = 0.0; pr.pop = 0; pr.node = -1; pr.valid = 0; pr.ntj0 = 0; pr.ntj1 = 0;
  if (*t.root >= d.n) {
    scaleAll(t, c);
    pr.valid = 1;
  }
  sd.prop[l] = pr;
}

// block partial sums: [0] data delta, [1] genealogy delta, [2] ntj0, [3] ntj1, [4..4+Q) coal totals, [4+Q..4+2Q) ncoal totals
// mode 0: deltas of the pending global proposal; mode 1: sum of data lnL in [0] and the statistics totals
__global__ void __launch_bounds__(kSmpThreads) k_smp_reduce(StoreDev d, SmpDev sd, int mode) {
  __shared__ double sh[kSmpThreads];
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int Q = sd.Q;
  const int V = 4 + 2 * Q;
  for (int v = 0; v < V; v++) {
    double x = 0.0;
    if (l < d.L) {
      if (mode == 0) {
        const SmpProposal pr = sd.prop[l];
        if (v == 0) x = pr.valid ? d.lnL[l] - d.savedLnL[l] : 0.0;
        else if (v == 1) x = pr.genDelta;
        else if (v == 2) x = pr.ntj0;
        else if (v == 3) x = pr.ntj1;
      } else {
        if (v == 0) x = d.lnL[l];
        else if (v >= 4 && v < 4 + Q) x = sd.coal[(size_t)l * Q + (v - 4)];
        else if (v >= 4 + Q) x = sd.ncoal[(size_t)l * Q + (v - 4 - Q)];
      }
    }
    if (mode == 0 && v >= 4) break;
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

// one global accept / reject for every locus; how: 0 tau move (statistics <- tentative), 1 rescaling (statistics *= c)
__global__ void __launch_bounds__(kSmpThreads) k_smp_global_resolve(StoreDev d, SmpDev sd, int accept, int how, double c) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= d.L) return;
  if (!sd.prop[l].valid) return;
  const TreeView t = deviceView(d, l);
  if (accept) {
    commit(t);
    double* c0 = sd.coal + (size_t)l * sd.Q;
    if (how == 0) {
      const double* cT = sd.coalT + (size_t)l * sd.Q;
      for (int p = 0; p < sd.Q; p++) c0[p] = cT[p];
    } else {
      for (int p = 0; p < sd.Q; p++) c0[p] *= c;
    }
  } else {
    revert(t);
  }
}

// statistics of every locus from scratch (initialisation and consistency checks)
__global__ void __launch_bounds__(kSmpThreads) k_smp_init_stats(StoreDev d, SmpDev sd, const SmpModel* __restrict__ mp, int toTentative,
                                                                int* __restrict__ ncoalOut) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= d.L) return;
  const SmpModel& m = *mp;
  const TreeView t = deviceView(d, l);
  if (*t.root < d.n) return;
  double* dst = (toTentative ? sd.coalT : sd.coal) + (size_t)l * m.Q;
  int* nc = (toTentative ? ncoalOut : sd.ncoal) + (size_t)l * m.Q;
  smpLocusStats(m, t.age, sd.nodePop + (size_t)l * d.N, -1, 0.0, dst, nc);
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
      if (!((m.below[q] >> np[r.left]) & 1ull) || !((m.below[q] >> np[r.right]) & 1ull)) v++;
      if (r.father >= 0 ? t.node[r.father].left != x && t.node[r.father].right != x : *t.root != x) v++;
      if (t.node[x].flags & (F_RECALC | F_SAVED)) v++;
    }
  }
  bad[l] = v;
}

}  // namespace gphocs
