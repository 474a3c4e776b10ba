// gen_kernels.cuh — genealogy likelihood P(G|M): coal/mig statistics + closed-form log-density.
//
// Replaces, for ALL loci in one launch, computeGenetreeStats (patch.c:2330-2354), recalcStats
// (patch.c:2387-2513), gtreeLnLikelihood (patch.c:2702-2723) and computeTotalStats (patch.c:2134-2164).
//
// Input is a flattened snapshot of the host's per-population event chains (patch.h:159-172):
//   evStart[l] .. evStart[l+1]          events of locus l (CSR)
//   popStart[l][p] .. popStart[l][p+1]   chain of population p, relative to evStart[l]
//   evTime[e]  fp64 elapsed_time,  evCode[e] = type | id << 3  (id = migration band where relevant)
// i.e. 10 bytes per event instead of the reference's 32-byte linked-list node.
//
// A CTA stages the events of a tile of 32 loci in shared memory with coalesced loads, then works per
// (population, locus) chain with the loci of the tile as the lanes of a warp: pass A sums the lineage deltas of
// each chain, a per-locus post-order sweep over the population tree turns them into lineages entering each
// population, pass B re-walks each chain in the reference's event order accumulating n(n-1)t and n*t per live
// band and leaves the chain's term of the log-density.  Per-population sums are sequential in chain order and
// products are not fused, so the statistics and the log-density are bit-identical to the reference on the same
// elapsed times.  Totals are reduced per CTA and then by a
// fixed-order second kernel (deterministic run to run).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gphocs {

constexpr int kGenTile = 32;      // loci per CTA
constexpr int kGenThreads = 128;
constexpr int kMaxPops = 39;      // 2*NSPECIES-1 (patch.h:19,49)
constexpr int kMaxBands = 100;    // MAX_MIG_BANDS  (patch.h:17)

enum { EV_COAL = 0, EV_IN_MIG, EV_OUT_MIG, EV_BAND_START, EV_BAND_END, EV_SAMPLES_START, EV_END_CHAIN, EV_DUMMY };

struct GenParams {
  int Q, C, B, rootPop;
  int postOrder[kMaxPops];
  int son0[kMaxPops], son1[kMaxPops];
  int samplesPerPop[kMaxPops];
  double theta[kMaxPops];
  double log2OverTheta[kMaxPops];  // log(2/theta), evaluated on the host once per parameter update
  double migRate[kMaxBands];
  double logMigRate[kMaxBands];
};

struct GenDev {
  int L, Q, B;
  const int* evStart;        // [L+1]
  const uint16_t* popStart;  // [L][Q+1]
  const double* evTime;
  const uint16_t* evCode;
  uint8_t* evLineages;       // optional [total events]
  double* lnL;               // [L]
  double* coal;              // [L][Q]
  int* numCoals;             // [L][Q]
  double* mig;               // [L][B]
  int* numMigs;              // [L][B]
  uint8_t* enter;            // [L][Q] lineages at the start of each chain (events[first].num_lineages, patch.c:2398)
  double* ctaTotals;         // [numCTAs][V], V = 1 + 2Q + 2B
  const GenParams* params;
};

__host__ __device__ inline int genTotalsLen(int Q, int B) { return 1 + 2 * Q + 2 * B; }

// lineages gained by an event of the given type, SAMPLES_START aside: COAL -1, IN_MIG -1, OUT_MIG +1, others 0
// (recalcStats' switch, patch.c:2415-2484) — 4-bit two's complement entries of a table in a register
__device__ __forceinline__ int lineageStep(int type) { return ((int)(0x000001ffu << (28 - 4 * type))) >> 28; }

__global__ void __launch_bounds__(kGenThreads) k_gen_eval(GenDev d, int maxTileEvents, const __grid_constant__ GenParams prm) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x;
  const int Q = d.Q, B = d.B, V = genTotalsLen(Q, B);
  const int l0 = blockIdx.x * kGenTile;
  const int nl = min(kGenTile, d.L - l0);
  // shared carve-up
  double* sTime = reinterpret_cast<double*>(smem);                     // [maxTileEvents]
  double* sMig = sTime + maxTileEvents;                                // [tile][B]
  double* sCoal = sMig + kGenTile * B;                                 // [tile][Q]
  double* sTermP = sCoal + kGenTile * Q;                               // [tile][Q] chain terms of the log-density
  double* sTot = sTermP + kGenTile * Q;                                // [V]
  double* sLnL = sTot + V;                                             // [tile]
  int* sNumCoals = reinterpret_cast<int*>(sLnL + kGenTile);            // [tile][Q]
  int* sNumMigs = sNumCoals + kGenTile * Q;                            // [tile][B]
  int* sEvBase = sNumMigs + kGenTile * B;                              // [tile+1] event offsets relative to tile
  uint16_t* sCode = reinterpret_cast<uint16_t*>(sEvBase + kGenTile + 1);  // [maxTileEvents]
  uint16_t* sPopStart = sCode + maxTileEvents;                         // [tile][Q+1]
  int16_t* sDelta = reinterpret_cast<int16_t*>(sPopStart + kGenTile * (Q + 1));  // [tile][Q] lineage delta, then n at chain start
  uint8_t* sEnd = reinterpret_cast<uint8_t*>(sDelta + kGenTile * Q);   // [tile][Q] lineages at the end of each chain
  // the model (populations, bands, parameters: 2.9 KB) arrives in the kernel's parameter space: every index into it
  // is uniform over a warp, so reads are constant-cache broadcasts and no CTA spends a round trip staging it

  const int e0 = d.evStart[l0];
  const int tileEvents = d.evStart[l0 + nl] - e0;
  if (tid <= nl) sEvBase[tid] = d.evStart[l0 + tid] - e0;
  // staging: four loads in flight per thread and array before the first store (the trip count is not known to the
  // compiler, which would otherwise wait for every load in turn)
  {
    const double* __restrict__ gt = d.evTime + e0;
    const uint16_t* __restrict__ gc = d.evCode + e0;
    int i = tid;
    for (; i + 3 * kGenThreads < tileEvents; i += 4 * kGenThreads) {
      const double t0 = gt[i], t1 = gt[i + kGenThreads], t2 = gt[i + 2 * kGenThreads], t3 = gt[i + 3 * kGenThreads];
      const uint16_t c0 = gc[i], c1 = gc[i + kGenThreads], c2 = gc[i + 2 * kGenThreads], c3 = gc[i + 3 * kGenThreads];
      sTime[i] = t0; sTime[i + kGenThreads] = t1; sTime[i + 2 * kGenThreads] = t2; sTime[i + 3 * kGenThreads] = t3;
      sCode[i] = c0; sCode[i + kGenThreads] = c1; sCode[i + 2 * kGenThreads] = c2; sCode[i + 3 * kGenThreads] = c3;
    }
    for (; i < tileEvents; i += kGenThreads) {
      sTime[i] = gt[i];
      sCode[i] = gc[i];
    }
    const uint16_t* __restrict__ gp = d.popStart + (size_t)l0 * (Q + 1);
    const int np = nl * (Q + 1);
    i = tid;
    for (; i + 2 * kGenThreads < np; i += 3 * kGenThreads) {
      const uint16_t p0 = gp[i], p1 = gp[i + kGenThreads], p2 = gp[i + 2 * kGenThreads];
      sPopStart[i] = p0; sPopStart[i + kGenThreads] = p1; sPopStart[i + 2 * kGenThreads] = p2;
    }
    for (; i < np; i += kGenThreads) sPopStart[i] = gp[i];
  }
  for (int i = tid; i < kGenTile * B; i += kGenThreads) { sMig[i] = 0.0; sNumMigs[i] = 0; }
  for (int i = tid; i < V; i += kGenThreads) sTot[i] = 0.0;
  __syncthreads();

  // Work items are (population, locus) chains with the 32 loci of the tile as the lanes of a warp: chains of one
  // population are about equally long in every locus, so the lanes stay together; the warps take the populations
  // in turn.
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int kGenWarps = kGenThreads / 32;
  // pass A: net lineage change of each chain (SAMPLES_START +samples, COAL -1, IN_MIG -1, OUT_MIG +1); event types only
  if (lane < nl) {
    const int j = lane, eb = sEvBase[j];
    const uint16_t* ps = sPopStart + j * (Q + 1);
    for (int p = warp; p < Q; p += kGenWarps) {
      const int a = eb + ps[p], b = eb + ps[p + 1];
      const int smp = prm.samplesPerPop[p];
      int delta = 0;
      for (int e = a; e < b; e++) {
        const int type = sCode[e] & 7;
        delta += type == EV_SAMPLES_START ? smp : lineageStep(type);
      }
      sDelta[j * Q + p] = (int16_t)delta;
    }
  }
  __syncthreads();
  // lineages entering each population: post-order over the population tree (patch.c:2336-2347); the deltas are
  // replaced by the number of lineages at the start of the chain
  if (tid < nl) {
    int16_t* dl = sDelta + tid * Q;
    for (int i = 0; i < Q; i++) {
      const int p = prm.postOrder[i];
      int n0 = 0;
      if (p >= prm.C) {
        const int s0 = prm.son0[p], s1 = prm.son1[p];
        n0 = sEnd[tid * Q + s0] + sEnd[tid * Q + s1];
      }
      sEnd[tid * Q + p] = (uint8_t)(n0 + dl[p]);
      dl[p] = (int16_t)n0;
      d.enter[(size_t)(l0 + tid) * Q + p] = (uint8_t)n0;   // what an incremental re-walk of one chain starts from
    }
  }
  __syncthreads();
  // pass B: statistics per chain in the reference's order of operations (patch.c:2403-2486), and the chain's term of
  // the log-density (patch.c:2709-2712) — the division is the expensive part of it and is off the per-locus sum
  if (lane < nl) {
    const int j = lane, eb = sEvBase[j];
    const uint16_t* ps = sPopStart + j * (Q + 1);
    double* mg = sMig + j * B;
    int* nm = sNumMigs + j * B;
    const bool wantLineages = d.evLineages != nullptr;
    for (int p = warp; p < Q; p += kGenWarps) {
      int n = sDelta[j * Q + p];
      const int a = eb + ps[p], b = eb + ps[p + 1];
      int ncoal = 0;
      double coal = 0.0;
      unsigned long long live0 = 0ull, live1 = 0ull;  // live migration bands (ids 0..127)
      const int smp = prm.samplesPerPop[p];
      // the next event is requested while the current one is processed (the walk is a chain of dependent
      // operations; its loads need not be part of it)
      double tNext = 0.0;
      int codeNext = 0;
      if (a < b) { tNext = sTime[a]; codeNext = sCode[a]; }
      for (int e = a; e < b; e++) {
        const double t = tNext;
        const int code = codeNext, type = code & 7;
        const int en = min(e + 1, b - 1);
        tNext = sTime[en]; codeNext = sCode[en];
        if (wantLineages) d.evLineages[e0 + e] = (uint8_t)n;
        coal = __dadd_rn(coal, __dmul_rn((double)(n * (n - 1)), t));
        // everything that touches a migration band: statistics of the live bands, band start / end, migrations.
        // A band's events lie in its target population's chain, so no two threads of a locus share a band.
        if ((live0 | live1) != 0ull || (type >= EV_IN_MIG && type <= EV_BAND_END)) {
          const int id = code >> 3;
          if (live0 | live1) {
            const double nt = __dmul_rn((double)n, t);
            for (unsigned long long m = live0; m; m &= m - 1) { const int bnd = __ffsll((long long)m) - 1; mg[bnd] = __dadd_rn(mg[bnd], nt); }
            for (unsigned long long m = live1; m; m &= m - 1) { const int bnd = 64 + __ffsll((long long)m) - 1; mg[bnd] = __dadd_rn(mg[bnd], nt); }
          }
          if (type == EV_IN_MIG) nm[id]++;
          else if (type == EV_BAND_START) { if (id < 64) live0 |= 1ull << id; else live1 |= 1ull << (id - 64); mg[id] = 0.0; nm[id] = 0; }
          else if (type == EV_BAND_END) { if (id < 64) live0 &= ~(1ull << id); else live1 &= ~(1ull << (id - 64)); }
        }
        ncoal += type == EV_COAL;
        n += type == EV_SAMPLES_START ? smp : lineageStep(type);
      }
      sCoal[j * Q + p] = coal;
      sNumCoals[j * Q + p] = ncoal;
      sTermP[j * Q + p] = __dsub_rn(__dmul_rn((double)ncoal, prm.log2OverTheta[p]), __ddiv_rn(coal, prm.theta[p]));
    }
  }
  __syncthreads();
  // per-locus log-density (patch.c:2709-2723): the chains' terms in population order, then the bands'
  if (tid < nl) {
    const int j = tid;
    double lnLd = 0.0;
    for (int p = 0; p < Q; p++) lnLd = __dadd_rn(lnLd, sTermP[j * Q + p]);
    for (int bnd = 0; bnd < B; bnd++) {
      const double m = prm.migRate[bnd];
      if (m > 0.0) {
        const double term = __dsub_rn(__dmul_rn((double)sNumMigs[j * B + bnd], prm.logMigRate[bnd]),
                                      __dmul_rn(sMig[j * B + bnd], m));
        lnLd = __dadd_rn(lnLd, term);
      }
    }
    sLnL[j] = lnLd;
    d.lnL[l0 + j] = lnLd;
  }
  __syncthreads();
  for (int i = tid; i < nl * Q; i += kGenThreads) {
    d.coal[(size_t)l0 * Q + i] = sCoal[i];
    d.numCoals[(size_t)l0 * Q + i] = sNumCoals[i];
  }
  for (int i = tid; i < nl * B; i += kGenThreads) {
    d.mig[(size_t)l0 * B + i] = sMig[i];
    d.numMigs[(size_t)l0 * B + i] = sNumMigs[i];
  }
  __syncthreads();
  // per-CTA totals, fixed order over the tile's loci
  for (int v = tid; v < V; v += kGenThreads) {
    double acc = 0.0;
    if (v == 0) {
      for (int j = 0; j < nl; j++) acc += sLnL[j];
    } else if (v < 1 + Q) {
      const double* x = sCoal + (v - 1);
      for (int j = 0; j < nl; j++) acc += x[j * Q];
    } else if (v < 1 + 2 * Q) {
      const int* x = sNumCoals + (v - 1 - Q);
      for (int j = 0; j < nl; j++) acc += (double)x[j * Q];
    } else if (v < 1 + 2 * Q + B) {
      const double* x = sMig + (v - 1 - 2 * Q);
      for (int j = 0; j < nl; j++) acc += x[j * B];
    } else {
      const int* x = sNumMigs + (v - 1 - 2 * Q - B);
      for (int j = 0; j < nl; j++) acc += (double)x[j * B];
    }
    d.ctaTotals[(size_t)blockIdx.x * V + v] = acc;
  }
}

// recalcStats (patch.c:2387-2513) for a list of (locus, population) chains whose events kept their order and number
// but changed their elapsed times — what rubberBand (patch.c:596-801) does to the chains of a split-time proposal.
// One thread per chain: the new times are copied over the old ones, the chain is re-walked in the reference's order of
// operations from the lineage count it is entered with (same code path as pass B of k_gen_eval, so the statistics are
// bit-identical to a full evaluation of the updated snapshot), the stored statistics are replaced and the change of the
// log-density is returned as recalcStats returns it: minus (mig - old) * rate at every MIG_BAND_END, in chain order,
// then minus (coal - old) / theta.  status[k] != 0: the chain has a different number of events (nothing is changed).
__global__ void __launch_bounds__(128) k_gen_recalc(GenDev d, double* __restrict__ evTime, int nPairs, const int* __restrict__ pairLocus,
                                                    const int* __restrict__ pairPop, const int* __restrict__ timesStart,
                                                    const double* __restrict__ newTimes, double* __restrict__ delta,
                                                    int* __restrict__ status, const __grid_constant__ GenParams prm) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nPairs) return;
  const int l = pairLocus[k], p = pairPop[k], Q = d.Q, B = d.B;
  if (l < 0 || l >= d.L || p < 0 || p >= Q) { status[k] = 3; delta[k] = 0.0; return; }
  const uint16_t* ps = d.popStart + (size_t)l * (Q + 1);
  const int a = d.evStart[l] + ps[p], b = d.evStart[l] + ps[p + 1];
  const int t0 = timesStart[k];
  if (timesStart[k + 1] - t0 != b - a) { status[k] = 1; delta[k] = 0.0; return; }
  status[k] = 0;
  int n = d.enter[(size_t)l * Q + p];
  const int smp = prm.samplesPerPop[p];
  double coal = 0.0, dl = 0.0;
  double* mg = d.mig + (size_t)l * B;
  double acc[8];             // statistics of the bands that are live at this point of the chain (they may overlap)
  int accId[8], nLive = 0;
  for (int e = a; e < b; e++) {
    const double t = newTimes[t0 + (e - a)];
    evTime[e] = t;
    const int code = d.evCode[e], type = code & 7, id = code >> 3;
    coal = __dadd_rn(coal, __dmul_rn((double)(n * (n - 1)), t));
    if (nLive) {
      const double nt = __dmul_rn((double)n, t);
      for (int i = 0; i < nLive; i++) acc[i] = __dadd_rn(acc[i], nt);
    }
    if (type == EV_BAND_START) {
      if (nLive < 8) { acc[nLive] = 0.0; accId[nLive] = id; nLive++; } else status[k] = 2;
    } else if (type == EV_BAND_END) {
      for (int i = 0; i < nLive; i++)
        if (accId[i] == id) {
          dl = __dsub_rn(dl, __dmul_rn(__dsub_rn(acc[i], mg[id]), prm.migRate[id]));
          mg[id] = acc[i];
          acc[i] = acc[nLive - 1]; accId[i] = accId[nLive - 1]; nLive--;
          break;
        }
    }
    n += type == EV_SAMPLES_START ? smp : lineageStep(type);
  }
  double* cs = d.coal + (size_t)l * Q + p;
  dl = __dsub_rn(dl, __ddiv_rn(__dsub_rn(coal, *cs), prm.theta[p]));
  *cs = coal;
  delta[k] = dl;
}

// event snapshot copied from page-locked caller arrays as it is: int32 type / id -> 16-bit code, int32 chain offsets
// -> 16-bit, int64 event offsets -> int32 relative to the first event; *bad counts malformed entries
__global__ void __launch_bounds__(256) k_gen_pack(const int* __restrict__ evType, const int* __restrict__ evId, long long numEvents,
                                                  const int* __restrict__ popStart, long long numPopStart,
                                                  const long long* __restrict__ evStart, int L, int B,
                                                  uint16_t* __restrict__ code, uint16_t* __restrict__ ps, int* __restrict__ es,
                                                  int* __restrict__ bad) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int wrong = 0;
  for (long long e = first; e < numEvents; e += stride) {
    const int t = evType[e], id = evId[e];
    const bool needsBand = (t == EV_IN_MIG || t == EV_BAND_START || t == EV_BAND_END);
    if (t < 0 || t > EV_DUMMY || (needsBand && (id < 0 || id >= B))) wrong = 1;
    code[e] = (uint16_t)(t | ((needsBand ? id : 0) << 3));
  }
  for (long long i = first; i < numPopStart; i += stride) ps[i] = (uint16_t)popStart[i];
  const long long e0 = evStart[0];
  for (long long l = first; l <= L; l += stride) {
    es[l] = (int)(evStart[l] - e0);
    if (l < L) {
      const long long nEv = evStart[l + 1] - evStart[l];
      if (nEv > 65535 || nEv < 0) wrong = 1;
    }
  }
  if (wrong) atomicAdd(bad, 1);
}

// a snapshot that arrived in the device's own format: the same checks k_gen_pack makes while it converts
__global__ void __launch_bounds__(256) k_gen_check_packed(const uint16_t* __restrict__ code, long long numEvents,
                                                          const uint16_t* __restrict__ ps, const int* __restrict__ es, int L,
                                                          int Q, int B, int* __restrict__ bad) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int wrong = 0;
  for (long long e = first; e < numEvents; e += stride) {
    const int t = code[e] & 7, id = code[e] >> 3;
    const bool needsBand = (t == EV_IN_MIG || t == EV_BAND_START || t == EV_BAND_END);
    if (needsBand ? id >= B : id != 0) wrong = 1;
  }
  for (long long l = first; l < L; l += stride) {
    const long long nEv = (long long)es[l + 1] - es[l];
    if (nEv > 65535 || nEv < 0) wrong = 1;
    const uint16_t* p = ps + l * (Q + 1);
    if (p[0] != 0 || (long long)p[Q] != nEv) wrong = 1;
    for (int q = 0; q < Q; q++)
      if (p[q + 1] < p[q]) wrong = 1;
  }
  if (first == 0 && es[0] != 0) wrong = 1;
  if (wrong) atomicAdd(bad, 1);
}

// out[v] = sum over CTAs of ctaTotals[cta][v], fixed order (one block per v)
__global__ void __launch_bounds__(256) k_gen_reduce(const double* __restrict__ ctaTotals, int numCtas, int V,
                                                    double* __restrict__ out) {
  __shared__ double sh[256];
  const int v = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < numCtas; i += 256) acc += ctaTotals[(size_t)i * V + v];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[v] = sh[0];
}

__host__ __device__ inline size_t genSmemBytes(int Q, int B, int maxTileEvents) {
  const int V = genTotalsLen(Q, B);
  size_t bytes = (size_t)maxTileEvents * 8 + (size_t)kGenTile * B * 8 + (size_t)kGenTile * Q * 8 * 2 + (size_t)V * 8 +
                 (size_t)kGenTile * 8;
  bytes += (size_t)kGenTile * Q * 4 + (size_t)kGenTile * B * 4 + (size_t)(kGenTile + 1) * 4;
  bytes += (size_t)maxTileEvents * 2 + (size_t)kGenTile * (Q + 1) * 2 + (size_t)kGenTile * Q * 3;
  return bytes + 32;
}

}  // namespace gphocs
