// ingest_kernels.cuh — alignment columns -> what initializeLocusData consumes:
// JC-canonical site patterns in order of first appearance with multiplicities (processLocusAlignment +
// cannonizeJCpattern, AlignmentProcessor.c:871-990, 1595-1655), the greedy choice of het genotypes that may be phased
// arbitrarily (computeHetSymmetryBreaks, :1706-1895) and all phasings of the others (processHetPatterns +
// getAllPhases, :998-1158, 2242-2290).
//
// One CTA per locus.  The kernel reads the TEXT of the sequence file (uploaded as it is; the host only locates the
// rows): characters are mapped to the canonical alphabet and validated here.  Columns are canonised by the threads
// (one column each, rows read coalesced), packed 4 bits per slot and de-duplicated in a shared-memory hash table that
// remembers first site and multiplicity (one atomic per distinct pattern per warp); the table is then ordered by
// first site.  The kernel runs twice: a counting pass (patterns U, phased columns P per locus), a
// host prefix sum, and an emitting pass that writes the final arrays — no per-locus scratch in HBM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gphocs {

constexpr int kIngThreads = 128;
constexpr int kIngColsPerThread = 8;                                // columns of one chunk held in registers
constexpr int kIngChunk = kIngThreads * kIngColsPerThread;
constexpr int kIngMaxSlots = 64;                                    // haploid slots (2 per diploid sample)
constexpr int kIngMaxHets = 32;                                     // diploid samples
constexpr int kIngMaxFree = 20;                                     // un-phased genotypes per pattern: 2^20 columns

// status bits per locus
enum { ING_OVERFLOW = 1, ING_COLLISION = 2, ING_TOO_MANY_PHASES = 4, ING_BAD_CHAR = 8 };

struct IngestTables {
  uint32_t permMask[15][15];   // permMask[symbol][image] = base permutations (24 bits) mapping symbol to image
  int8_t slotRow[kIngMaxSlots];    // row (named sample) of the slot, -1: second slot of a diploid (all N)
  uint8_t isDiploid[kIngMaxSlots];
  int8_t symbolOf[256];            // character -> index in "TCAGYWKMSRVDBHN" (either case), -1: not a base symbol
};

struct IngestDev {
  int L, n, R;                       // loci, slots, rows (named samples)
  const char* text;                  // the sequence file
  const long long* rowOff;           // [L][R] offset of the first base of the row in `text`, -1: sample absent (all N)
  const int* seqLen;                 // [L]
  unsigned long long* firstBad;      // out: smallest (locus << 40 | row << 32 | site) holding an illegal character
  const int* locusIds;               // NULL: all loci; else the loci this launch handles (grid = count)
  // pattern mode (processHetPatterns called on its own): canonical patterns + counts instead of raw columns
  const uint8_t* givenPatterns;      // [sum U][n] symbol indices, locus l at givenStart[l]
  const int* givenCounts;
  const int* givenStart;             // [L+1]
  int breakSymmetries;
  int* numPatterns;                  // [L] out (count pass)
  int* numPhased;                    // [L] out (count pass)
  int* status;                       // [L] out
  // emit pass
  const int* pattStart;              // [L+1] phased columns
  const int* unphStart;              // [L+1]
  char* chars;                       // [sum P][n]
  int* numPhases;                    // [sum P]
  int* counts;                       // [sum U]
  uint8_t* canon;                    // [sum U][n] canonical pattern characters (may be NULL)
};

// shared-memory carve-up for a table of H slots (at most H/2 distinct patterns), W 64-bit words per key
struct IngestSmem {
  int tag, first, cnt, fullKey, list, okey, ocnt, brk, score, nhets, where, live, pstart, total;
};
__host__ __device__ inline IngestSmem ingestSmemLayout(int H, int W) {
  IngestSmem s;
  const int U = H / 2;
  int o = 0;
  s.tag = o;      o += 8 * H;        // u64 tags      \  after ordering this region (16 H bytes) is reused for the
  s.first = o;    o += 4 * H;        // first site    |  het lists: 32 bytes per distinct pattern
  s.cnt = o;      o += 4 * H;        // multiplicity  /
  s.fullKey = o;  o += 8 * W * H;
  s.okey = o;     o += 8 * W * U;    // keys ordered by first site
  s.brk = o;      o += 8 * U;        // arbitrarily phased slots (bit mask)
  s.score = o;    o += 8 * U;
  s.ocnt = o;     o += 4 * U;
  s.pstart = o;   o += 4 * (U + 1);
  s.list = o;     o += 2 * U;
  s.where = o;    o += 2 * U;
  s.live = o;     o += 2 * U;
  s.nhets = o;    o += U;
  s.total = (o + 15) & ~15;
  return s;
}

__device__ inline uint64_t ingMix(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

__device__ inline int ingSymbolOf(const uint64_t* key, int slot) { return (int)((key[slot >> 4] >> ((slot & 15) * 4)) & 15u); }
__device__ inline bool ingPartial(int sym) { return sym >= 4 && sym < 14; }   // two- and three-way codes

// the two bases of a diploid genotype (translateAmbiguity, :2298-2340): index into "TCAGN"
__device__ inline void ingGenotype(int sym, int* a, int* b) {
  // symbols: T C A G | Y(TC) W(TA) K(TG) M(AC) S(CG) R(AG) | V D B H | N
  const int first[15] = {0, 1, 2, 3, 0, 0, 0, 2, 1, 2, 4, 4, 4, 4, 4};
  const int second[15] = {0, 1, 2, 3, 1, 2, 3, 1, 3, 3, 4, 4, 4, 4, 4};
  *a = first[sym];
  *b = second[sym];
}

template <int W>
__global__ void __launch_bounds__(kIngThreads)
k_ingest(IngestDev d, const IngestTables* __restrict__ tabs, int H, int emit, uint64_t salt) {
  extern __shared__ __align__(16) unsigned char ingSmem[];
  __shared__ uint32_t sPerm[15][15];
  __shared__ int8_t sSym[256];
  __shared__ long long sRowOff[kIngMaxSlots];   // per slot: offset of the row in the text, -1 = all N
  __shared__ uint8_t sHaploid[kIngMaxSlots];
  __shared__ int sNumUnique, sFlags, sTotal;
  const int tid = threadIdx.x;
  const int l = d.locusIds ? d.locusIds[blockIdx.x] : blockIdx.x;
  const int n = d.n;
  const IngestSmem lay = ingestSmemLayout(H, W);
  unsigned long long* tag = (unsigned long long*)(ingSmem + lay.tag);
  int* first = (int*)(ingSmem + lay.first);
  int* cnt = (int*)(ingSmem + lay.cnt);
  uint64_t* fullKey = (uint64_t*)(ingSmem + lay.fullKey);
  uint64_t* okey = (uint64_t*)(ingSmem + lay.okey);
  uint64_t* brk = (uint64_t*)(ingSmem + lay.brk);
  long long* score = (long long*)(ingSmem + lay.score);
  int* ocnt = (int*)(ingSmem + lay.ocnt);
  int* pstart = (int*)(ingSmem + lay.pstart);
  uint16_t* list = (uint16_t*)(ingSmem + lay.list);
  int16_t* where = (int16_t*)(ingSmem + lay.where);
  uint16_t* live = (uint16_t*)(ingSmem + lay.live);
  uint8_t* nhets = (uint8_t*)(ingSmem + lay.nhets);
  uint8_t* hets = ingSmem + lay.tag;          // aliases tag/first/cnt once the patterns are ordered
  const int Umax = H / 2;

  for (int i = tid; i < 225; i += kIngThreads) sPerm[i / 15][i % 15] = tabs->permMask[i / 15][i % 15];
  for (int i = tid; i < 256; i += kIngThreads) sSym[i] = tabs->symbolOf[i];
  if (tid < n) {
    const int row = tabs->slotRow[tid];
    sRowOff[tid] = (row >= 0 && d.givenPatterns == nullptr) ? d.rowOff[(size_t)l * d.R + row] : -1;
    sHaploid[tid] = tabs->isDiploid[tid] ? 0 : 1;
  }
  for (int h = tid; h < H; h += kIngThreads) { tag[h] = 0ull; first[h] = 0x7fffffff; cnt[h] = 0; }
  if (tid == 0) { sNumUnique = 0; sFlags = 0; sTotal = 0; }
  __syncthreads();

  int U = 0;
  if (d.givenPatterns == nullptr) {
    // ---------------------------------------------------------------- columns -> distinct canonical patterns
    const int S = d.seqLen[l];
    for (int c0 = 0; c0 < S; c0 += kIngChunk) {
      uint64_t key[kIngColsPerThread][W];
      int slotOf[kIngColsPerThread];
#pragma unroll
      for (int i = 0; i < kIngColsPerThread; i++) {
        const int site = c0 + i * kIngThreads + tid;
        slotOf[i] = -1;
        bool informative = false;
#pragma unroll
        for (int w = 0; w < W; w++) key[i][w] = 0ull;
        if (site < S) {
          uint32_t alive = 0xFFFFFFu;
          for (int s = 0; s < n; s++) {
            const long long off = sRowOff[s];
            int sym = 14;
            if (off >= 0) {
              sym = sSym[(unsigned char)d.text[off + site]];
              // readSeqs (:806-826): not a base symbol, or an ambiguity code in a haploid sample
              if (sym < 0 || (sym >= 4 && sym < 14 && sHaploid[s])) {
                atomicMin(d.firstBad, (unsigned long long)l << 40 | (unsigned long long)tabs->slotRow[s] << 32 | (unsigned)site);
                atomicOr(&sFlags, ING_BAD_CHAR);
                sym = 14;
              }
            }
            int image = 14;
            if (sym != 14) {
              informative = true;
              const int lo = sym < 4 ? 0 : (sym < 10 ? 4 : 10), hi = sym < 4 ? 3 : (sym < 10 ? 9 : 13);
              for (image = lo; image < hi; image++)
                if (alive & sPerm[sym][image]) break;
              alive &= sPerm[sym][image];
            }
            key[i][s >> 4] |= (uint64_t)image << ((s & 15) * 4);
          }
        }
        // columns of N only are dropped (:910-912).  One lane per distinct pattern of the warp does the insertion.
        uint64_t hsh = salt;
#pragma unroll
        for (int w = 0; w < W; w++) hsh = ingMix(hsh ^ key[i][w]);
        const unsigned long long t = informative ? (hsh | 1ull) : 0ull;
        const unsigned peers = __match_any_sync(0xffffffffu, t);
        const int leader = __ffs(peers) - 1;
        int h = -1;
        if (informative && (tid & 31) == leader) {
          h = (int)(hsh >> 20) & (H - 1);
          int probes = 0;
          for (;; h = (h + 1) & (H - 1)) {
            const unsigned long long old = atomicCAS(&tag[h], 0ull, t);
            if (old == 0ull) {
              if (atomicAdd(&sNumUnique, 1) >= Umax) atomicOr(&sFlags, ING_OVERFLOW);
              break;
            }
            if (old == t) break;
            if (++probes >= H) { atomicOr(&sFlags, ING_OVERFLOW); h = -1; break; }
          }
          if (h >= 0) {
            atomicMin(&first[h], site);              // the leader is the lowest lane = the lowest site of its peers
            atomicAdd(&cnt[h], __popc(peers));
          }
        }
        h = __shfl_sync(0xffffffffu, h, leader);
        slotOf[i] = informative ? h : -1;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kIngColsPerThread; i++)
        if (slotOf[i] >= 0 && first[slotOf[i]] == c0 + i * kIngThreads + tid) {
#pragma unroll
          for (int w = 0; w < W; w++) fullKey[(size_t)slotOf[i] * W + w] = key[i][w];
        }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kIngColsPerThread; i++)
        if (slotOf[i] >= 0) {
          bool same = true;
#pragma unroll
          for (int w = 0; w < W; w++) same &= fullKey[(size_t)slotOf[i] * W + w] == key[i][w];
          if (!same) atomicOr(&sFlags, ING_COLLISION);   // two patterns with one 64-bit tag: the host re-salts
        }
    }
    __syncthreads();
    if (sFlags) {
      if (tid == 0) { d.status[l] = sFlags; if (!emit) { d.numPatterns[l] = 0; d.numPhased[l] = 0; } }
      return;
    }
    U = sNumUnique;
    __syncthreads();
    if (tid == 0) sNumUnique = 0;
    __syncthreads();
    for (int h = tid; h < H; h += kIngThreads)
      if (tag[h] != 0ull) list[atomicAdd(&sNumUnique, 1)] = (uint16_t)h;
    __syncthreads();
    for (int e = tid; e < U; e += kIngThreads) {   // order of first appearance
      const int mine = first[list[e]];
      int rank = 0;
      for (int j = 0; j < U; j++) rank += first[list[j]] < mine;
      const int h = list[e];
#pragma unroll
      for (int w = 0; w < W; w++) okey[(size_t)rank * W + w] = fullKey[(size_t)h * W + w];
      ocnt[rank] = cnt[h];
    }
  } else {
    // ---------------------------------------------------------------- patterns given (processHetPatterns alone)
    const int g0 = d.givenStart[l];
    U = d.givenStart[l + 1] - g0;
    if (U > Umax) {
      if (tid == 0) { d.status[l] = ING_OVERFLOW; if (!emit) { d.numPatterns[l] = 0; d.numPhased[l] = 0; } }
      return;
    }
    for (int u = tid; u < U; u += kIngThreads) {
      uint64_t k[W];
#pragma unroll
      for (int w = 0; w < W; w++) k[w] = 0ull;
      for (int s = 0; s < n; s++) k[s >> 4] |= (uint64_t)d.givenPatterns[(size_t)(g0 + u) * n + s] << ((s & 15) * 4);
#pragma unroll
      for (int w = 0; w < W; w++) okey[(size_t)u * W + w] = k[w];
      ocnt[u] = d.givenCounts[g0 + u];
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ het genotypes that may be phased arbitrarily
  // het lists (slot ids in slot order) of the eligible patterns: seen once, at least one partial ambiguity
  for (int u = tid; u < U; u += kIngThreads) {
    int k = 0;
    brk[u] = 0ull;
    if (ocnt[u] <= 1)
      for (int s = 0; s < n; s++)
        if (ingPartial(ingSymbolOf(okey + (size_t)u * W, s)) && k < kIngMaxHets) hets[(size_t)u * kIngMaxHets + k++] = (uint8_t)s;
    nhets[u] = (uint8_t)k;
    score[u] = k > 0 ? (1ll << k) : -1ll;
    where[u] = -1;
  }
  __syncthreads();
  if (tid == 0 && d.breakSymmetries) {
    int numLive = 0, chosen = -1;
    long long top = -1;
    for (int u = 0; u < U; u++) {
      if (nhets[u] > 0) { where[u] = (int16_t)numLive; live[numLive++] = (uint16_t)u; }
      if (top < score[u]) { top = score[u]; chosen = u; }
    }
    while (top > 0) {
      const int slot = hets[(size_t)chosen * kIngMaxHets + --nhets[chosen]];
      brk[chosen] |= 1ull << slot;
      score[chosen] = nhets[chosen] == 0 ? -1ll : score[chosen] / 2;
      top = score[chosen];
      for (int i = 0; i < numLive;) {
        const int u = live[i];
        uint8_t* h = hets + (size_t)u * kIngMaxHets;
        for (int k = 0; k < nhets[u]; k++)
          if (h[k] == slot) { h[k] = h[--nhets[u]]; break; }
        if (nhets[u] > 0) {
          i++;
        } else {
          numLive--;
          live[where[u]] = live[numLive];
          where[live[where[u]]] = where[u];
          where[u] = -1;
          score[u] = -1ll;
        }
        if (top < score[u]) { top = score[u]; chosen = u; }
      }
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ phased columns per pattern
  for (int u = tid; u < U; u += kIngThreads) {
    int numFree = 0;
    for (int s = 0; s < n; s++)
      if (ingPartial(ingSymbolOf(okey + (size_t)u * W, s)) && !((brk[u] >> s) & 1ull)) numFree++;
    if (numFree > kIngMaxFree) { atomicOr(&sFlags, ING_TOO_MANY_PHASES); numFree = 0; }
    pstart[u + 1] = 1 << numFree;
  }
  __syncthreads();
  if (tid == 0) {
    pstart[0] = 0;
    long long run = 0;
    for (int u = 0; u < U; u++) {
      run += pstart[u + 1];
      if (run > 0x3fffffff) { sFlags |= ING_TOO_MANY_PHASES; run = 0; }
      pstart[u + 1] = (int)run;
    }
    sTotal = (int)run;
  }
  __syncthreads();
  if (!emit) {
    if (tid == 0) { d.numPatterns[l] = sFlags ? 0 : U; d.numPhased[l] = sFlags ? 0 : sTotal; d.status[l] = sFlags; }
    return;
  }
  if (sFlags) return;

  // ------------------------------------------------------------------ emit
  const char sym2char[16] = {'T', 'C', 'A', 'G', 'Y', 'W', 'K', 'M', 'S', 'R', 'V', 'D', 'B', 'H', 'N', '?'};
  const char base2char[5] = {'T', 'C', 'A', 'G', 'N'};
  const int P = sTotal;
  const int p0 = d.pattStart[l], u0 = d.unphStart[l];
  for (int u = tid; u < U; u += kIngThreads) {
    d.counts[u0 + u] = ocnt[u];
    if (d.canon)
      for (int s = 0; s < n; s++) d.canon[(size_t)(u0 + u) * n + s] = (uint8_t)sym2char[ingSymbolOf(okey + (size_t)u * W, s)];
  }
  for (int j = tid; j < P; j += kIngThreads) {
    int lo = 0, hi = U - 1;                 // pattern u with pstart[u] <= j < pstart[u+1]
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (pstart[mid] <= j) lo = mid; else hi = mid - 1;
    }
    const int u = lo, k = j - pstart[u];
    const uint64_t* key = okey + (size_t)u * W;
    char* col = d.chars + (size_t)(p0 + j) * n;
    int freeIdx = 0;
    for (int s = 0; s < n; s++) {
      const int sym = ingSymbolOf(key, s);
      if (!tabs->isDiploid[s]) { col[s] = sym2char[sym]; continue; }
      int a, b;
      ingGenotype(sym, &a, &b);
      // column k swaps the two bases of the j-th un-phased genotype iff bit j of k is set (getAllPhases counts the
      // phasings like a binary counter over the free genotypes in slot order, :2258-2283)
      if (ingPartial(sym) && !((brk[u] >> s) & 1ull)) {
        if ((k >> freeIdx) & 1) { const int t = a; a = b; b = t; }
        freeIdx++;
      }
      col[s] = base2char[a];
      col[s + 1] = base2char[b];
      s++;
    }
    d.numPhases[p0 + j] = k == 0 ? pstart[u + 1] - pstart[u] : 0;
  }
}

}  // namespace gphocs
