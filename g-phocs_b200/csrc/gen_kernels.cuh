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
// A CTA stages the events of a tile of 32 loci in shared memory with coalesced loads; then one thread per
// locus walks the populations in post-order (lineages entering an ancestral population = lineages its sons
// end with) and every chain in the reference's event order, accumulating n(n-1)t and n*t per live band.
// Per-population sums are sequential in chain order and products are not fused, so the statistics and the
// log-density are bit-identical to the reference on the same elapsed times.  Totals are reduced per CTA and then by a
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
  double* ctaTotals;         // [numCTAs][V], V = 1 + 2Q + 2B
  const GenParams* params;
};

__host__ __device__ inline int genTotalsLen(int Q, int B) { return 1 + 2 * Q + 2 * B; }

__global__ void __launch_bounds__(kGenThreads) k_gen_eval(GenDev d, int maxTileEvents) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x;
  const int Q = d.Q, B = d.B, V = genTotalsLen(Q, B);
  const int l0 = blockIdx.x * kGenTile;
  const int nl = min(kGenTile, d.L - l0);
  // shared carve-up
  double* sTime = reinterpret_cast<double*>(smem);                     // [maxTileEvents]
  double* sMig = sTime + maxTileEvents;                                // [tile][B]
  double* sCoal = sMig + kGenTile * B;                                 // [tile][Q]
  double* sTot = sCoal + kGenTile * Q;                                 // [V]
  double* sLnL = sTot + V;                                             // [tile]
  int* sDelta = reinterpret_cast<int*>(sLnL + kGenTile);               // [tile][Q] lineages at the end of each chain
  int* sNumCoals = sDelta + kGenTile * Q;                              // [tile][Q]
  int* sNumMigs = sNumCoals + kGenTile * Q;                            // [tile][B]
  int* sEvBase = sNumMigs + kGenTile * B;                              // [tile+1] event offsets relative to tile
  uint16_t* sCode = reinterpret_cast<uint16_t*>(sEvBase + kGenTile + 1);  // [maxTileEvents]
  uint16_t* sPopStart = sCode + maxTileEvents;                         // [tile][Q+1]
  __shared__ GenParams prm;

  for (int i = tid; i < (int)(sizeof(GenParams) / sizeof(int)); i += kGenThreads)
    reinterpret_cast<int*>(&prm)[i] = reinterpret_cast<const int*>(d.params)[i];
  const int e0 = d.evStart[l0];
  const int tileEvents = d.evStart[l0 + nl] - e0;
  if (tid <= nl) sEvBase[tid] = d.evStart[l0 + tid] - e0;
  for (int i = tid; i < tileEvents; i += kGenThreads) {
    sTime[i] = d.evTime[e0 + i];
    sCode[i] = d.evCode[e0 + i];
  }
  for (int i = tid; i < nl * (Q + 1); i += kGenThreads) sPopStart[i] = d.popStart[(size_t)l0 * (Q + 1) + i];
  for (int i = tid; i < kGenTile * B; i += kGenThreads) { sMig[i] = 0.0; sNumMigs[i] = 0; }
  for (int i = tid; i < V; i += kGenThreads) sTot[i] = 0.0;
  __syncthreads();

  // One thread per locus walks the whole genealogy: populations in post-order (patch.c:2336-2347: an ancestral
  // population starts with the lineages its two sons end with, a leaf population with none), every chain in the
  // reference's event order accumulating n(n-1)t and n*t per live band (patch.c:2403-2486), then the log-density
  // (patch.c:2709-2723).  The 32 loci of the tile are the lanes of one warp: chains of a population are about
  // equally long in every locus, so the lanes stay together, and the whole tile costs one chain walk instead of
  // one per (locus, population).
  if (tid < nl) {
    const int j = tid;
    const int eb = sEvBase[j];
    const uint16_t* ps = sPopStart + j * (Q + 1);
    int* nEnd = sDelta + j * Q;        // lineages at the end of each chain
    double* mg = sMig + j * B;
    int* nm = sNumMigs + j * B;
    const bool wantLineages = d.evLineages != nullptr;
    for (int i = 0; i < Q; i++) {
      const int p = prm.postOrder[i];
      int n = p >= prm.C ? nEnd[prm.son0[p]] + nEnd[prm.son1[p]] : 0;
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
        // everything that touches a migration band: statistics of the live bands, band start / end, migrations
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
        n += type == EV_SAMPLES_START ? smp : (type == EV_COAL || type == EV_IN_MIG) ? -1 : type == EV_OUT_MIG ? 1 : 0;
      }
      nEnd[p] = n;
      sCoal[j * Q + p] = coal;
      sNumCoals[j * Q + p] = ncoal;
    }
    // per-locus log-density (patch.c:2709-2723)
    double lnLd = 0.0;
    for (int p = 0; p < Q; p++) {
      const double term = __dsub_rn(__dmul_rn((double)sNumCoals[j * Q + p], prm.log2OverTheta[p]),
                                    __ddiv_rn(sCoal[j * Q + p], prm.theta[p]));
      lnLd = __dadd_rn(lnLd, term);
    }
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
    for (int j = 0; j < nl; j++) {
      double x;
      if (v == 0) x = sLnL[j];
      else if (v < 1 + Q) x = sCoal[j * Q + (v - 1)];
      else if (v < 1 + 2 * Q) x = (double)sNumCoals[j * Q + (v - 1 - Q)];
      else if (v < 1 + 2 * Q + B) x = sMig[j * B + (v - 1 - 2 * Q)];
      else x = (double)sNumMigs[j * B + (v - 1 - 2 * Q - B)];
      acc += x;
    }
    d.ctaTotals[(size_t)blockIdx.x * V + v] = acc;
  }
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
  size_t bytes = (size_t)maxTileEvents * 8 + (size_t)kGenTile * B * 8 + (size_t)kGenTile * Q * 8 + (size_t)V * 8 +
                 (size_t)kGenTile * 8;
  bytes += (size_t)kGenTile * Q * 4 * 2 + (size_t)kGenTile * B * 4 + (size_t)(kGenTile + 1) * 4;
  bytes += (size_t)maxTileEvents * 2 + (size_t)kGenTile * (Q + 1) * 2;
  return bytes + 32;
}

}  // namespace gphocs
