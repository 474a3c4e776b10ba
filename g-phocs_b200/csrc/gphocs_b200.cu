// gphocs_b200.cu — libgphocs_b200.so: device-resident locus store, batched engine, genealogy likelihood
// and the reference-compatible LocusData C ABI (include/gphocs_b200.h).  sm_100a only, no CPU fallback:
// every entry point that evaluates a likelihood launches a CUDA kernel or fails loudly.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <vector>

#include "gphocs_b200.h"
#include "host_runtime.h"

#include "clv_kernels.cuh"
#include "gen_kernels.cuh"
#include "ingest_kernels.cuh"
#include "sampler_kernels.cuh"
#include "sampler_mig.cuh"
#include "sweep_kernels.cuh"
#include "tree_ops.cuh"

using namespace gphocs;

static_assert(sizeof(GphocsOp) == sizeof(Op), "edit record layout");

static std::atomic<long long> g_launches{0};

#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) {                                                                             \
      fprintf(stderr, "gphocs_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e), __FILE__, __LINE__, \
              cudaGetErrorString(_e));                                                                   \
      return -1;                                                                                         \
    }                                                                                                    \
  } while (0)

template <typename T>
static int devAlloc(T** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  CUDA_TRY(cudaMalloc((void**)p, count * sizeof(T)));
  return 0;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is one setting per function and device, shared by every store /
// genealogy / sampler of the process: only ever raise it (per function: the largest request seen so far), so that a
// second object with a smaller need cannot lower the limit under the first one's next launch.
static int raiseDynamicSmem(const void* func, size_t bytes) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, size_t>> seen;
  std::lock_guard<std::mutex> lk(mu);
  size_t want = bytes;
  bool found = false;
  for (auto& e : seen)
    if (e.first == func) { found = true; e.second = want = std::max(e.second, bytes); }
  if (!found) seen.emplace_back(func, bytes);
  CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
  return 0;
}

// growable pinned host + device buffer pair
template <typename T>
struct Staging {
  T* host = nullptr;
  T* dev = nullptr;
  size_t cap = 0;
  int reserve(size_t n) {
    if (n <= cap) return 0;
    size_t nc = std::max(n, cap * 2);
    if (host) cudaFreeHost(host);
    if (dev) cudaFree(dev);
    host = nullptr; dev = nullptr; cap = 0;
    CUDA_TRY(cudaMallocHost((void**)&host, nc * sizeof(T)));
    CUDA_TRY(cudaMalloc((void**)&dev, nc * sizeof(T)));
    cap = nc;
    return 0;
  }
  void release() {
    if (host) cudaFreeHost(host);
    if (dev) cudaFree(dev);
    host = nullptr; dev = nullptr; cap = 0;
  }
};

// growable device-only buffer
template <typename T>
struct DevBuf {
  T* dev = nullptr;
  size_t cap = 0;
  int reserve(size_t n) {
    if (n <= cap) return 0;
    if (dev) cudaFree(dev);
    dev = nullptr; cap = 0;
    CUDA_TRY(cudaMalloc((void**)&dev, n * sizeof(T)));
    cap = n;
    return 0;
  }
  void release() {
    if (dev) cudaFree(dev);
    dev = nullptr; cap = 0;
  }
};

// ======================================================================================= store
struct GphocsStore {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  int L = 0, n = 0, N = 0, NI = 0, W = 0;
  long long Ct = 0;
  long long deviceBytes = 0;
  StoreDev d{};
  uint8_t* dMask = nullptr;
  Batch* dBatches = nullptr;
  bool anyOversized = false;   // some locus has more than kThreads columns (a CTA batch of its own)
  double* dSum = nullptr;
  int numBatches = 0;
  int maxBatchLoci = 1;  // most loci any CTA batch holds (sizes the kernel's shared memory)
  int prefetchAhead = 0; // CTAs resident on the whole GPU at once = how far ahead a CTA prefetches into L2 (incremental
                         // evaluations only: measured +2-3 % there, -2 % on a full evaluation, whose CTAs live long enough)
  size_t smemBytes = 0;
  std::vector<Batch> batches;
  std::vector<int> locusBatch;  // batch index holding each locus (-1: no columns)
  std::vector<int> colStart;
  // host mirror (what the reference's getters read)
  std::vector<NodeRec> hNode, hsNode;
  std::vector<double> hAge, hsAge, hRate, hLnL, hSavedLnL;
  std::vector<int> hRoot, hSavedRoot;
  // staging
  Staging<Op> ops;
  Staging<int> seg, status, ids;
  Staging<double> f64;
  Staging<int16_t> i16;
  int* dTopo32 = nullptr;      // device scratch for int32 topology copied straight from page-locked caller arrays
  size_t topo32Cap = 0;
  int* dBadTopo = nullptr;
  std::vector<Op> pending;  // edits queued by the scalar API, flushed before the next evaluation
  bool debugMirror = false;  // also mirror SEL/RECALC bits on the host (tests)
  bool opsInFlight = false;  // an edit batch was enqueued without a stream synchronisation
  std::atomic<bool> mirrorStale{false};  // topology / ages / roots of the mirror are behind the device copy (refreshMirrorLocked)
  std::atomic<bool> flagsStale{false};   // edits went to the device only (gphocsStoreApplyOpsAsync): the mirror's saved copies and
                                         // flag bytes follow with the next refresh as well
  Op* dOpsAsync = nullptr;               // device copy of the records of gphocsStoreApplyOpsAsync
  size_t opsAsyncCap = 0;
  int* dBadOps = nullptr;
  std::atomic<bool> lnlStale{false};     // lnL / savedLnL of the mirror are behind it (gphocsStoreEvaluateDevice, the sampler)
  std::mutex mu;

  TreeView hostView(int l) {
    TreeView t;
    const size_t o = (size_t)l * N;
    t.node = hNode.data() + o; t.saved = hsNode.data() + o;
    t.age = hAge.data() + o; t.svAge = hsAge.data() + o;
    t.root = &hRoot[l]; t.savedRoot = &hSavedRoot[l];
    t.lnL = &hLnL[l]; t.savedLnL = &hSavedLnL[l]; t.rate = &hRate[l];
    t.numLeaves = n;
    t.numPatterns = colStart[l + 1] - colStart[l];
    return t;
  }
};

static inline unsigned leafCode(char ch) {
  switch (ch) {  // computeLeafConditionals, LocusDataLikelihood.c:1336-1386
    case 'T': return 1u;
    case 'C': return 2u;
    case 'A': return 4u;
    case 'G': return 8u;
    case 'N': return 15u;
    default: return 0u;
  }
}

extern "C" GphocsStore* gphocsStoreCreate(int device, int numLoci, int numLeaves, const long long* pattStart,
                                          const long long* unphStart, const char* chars, const int* numPhases,
                                          const int* counts) {
  if (numLoci <= 0 || numLeaves < 2 || numLeaves > 200) {
    fprintf(stderr, "gphocs_b200: bad store dimensions (loci %d, leaves %d; leaves must be 2..200)\n", numLoci, numLeaves);
    return nullptr;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "gphocs_b200: no CUDA device visible — this library has no CPU path\n");
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    fprintf(stderr, "gphocs_b200: cannot select CUDA device %d\n", device);
    return nullptr;
  }
  GphocsStore* s = new GphocsStore();
  s->device = device;
  s->L = numLoci; s->n = numLeaves; s->N = 2 * numLeaves - 1; s->NI = numLeaves - 1; s->W = (numLeaves + 15) / 16;
  const int L = s->L, n = s->n, N = s->N, NI = s->NI, W = s->W;
  const long long Ct = pattStart[L] - pattStart[0];
  if (Ct >= (1ll << 31) - 256) {
    fprintf(stderr, "gphocs_b200: too many pattern columns (%lld)\n", Ct);
    delete s;
    return nullptr;
  }
  s->Ct = Ct;
  s->colStart.resize(L + 1);
  for (int l = 0; l <= L; l++) s->colStart[l] = (int)(pattStart[l] - pattStart[0]);

  // ---- host-side packing of leaves and phase groups
  std::vector<unsigned long long> words((size_t)W * std::max<long long>(Ct, 1), 0ull);
  std::vector<int> phases(std::max<long long>(Ct, 1), 0), cnt(std::max<long long>(Ct, 1), 0);
  bool bad = false;
  parallelFor(0, L, [&](long long lo_, long long hi_) {
  for (int l = (int)lo_; l < (int)hi_; l++) {
    long long u = unphStart[l];
    for (long long c = pattStart[l]; c < pattStart[l + 1]; c++) {
      const long long col = c - pattStart[0];
      const char* row = chars + (size_t)c * n;
      for (int leaf = 0; leaf < n; leaf++) {
        const unsigned code = leafCode(row[leaf]);
        if (!code) bad = true;
        words[(size_t)(leaf >> 4) * Ct + col] |= (unsigned long long)code << ((leaf & 15) * 4);
      }
      phases[col] = numPhases[c];
      if (numPhases[c] > 0) {
        cnt[col] = counts ? counts[u] : 0;
        u++;
        if (cnt[col] <= 0) bad = true;  // the reference aborts on a dead pattern (.c:450-453)
      }
    }
  }
  });
  if (bad) {
    fprintf(stderr, "gphocs_b200: unexpected character in a pattern (only T,C,A,G,N) or non-positive pattern count\n");
    delete s;
    return nullptr;
  }
  // phase groups must be whole and lie inside their locus
  for (int l = 0; l < L && !bad; l++)
    for (int c = s->colStart[l]; c < s->colStart[l + 1];) {
      const int ph = phases[c];
      if (ph <= 0 || c + ph > s->colStart[l + 1]) { bad = true; break; }
      for (int j = 1; j < ph; j++) if (phases[c + j] != 0) bad = true;
      c += ph;
    }
  if (bad) {
    fprintf(stderr, "gphocs_b200: malformed phase groups (numPhases must be 2^k on the first column, 0 on the rest)\n");
    delete s;
    return nullptr;
  }

  // ---- CTA batches: whole loci packed greedily; an oversized locus gets a CTA of its own
  long long scratchCols = 0;
  s->locusBatch.assign(L, -1);
  // loci per batch are also capped so that the per-locus staging area fits a 64 KB shared-memory budget
  // (keeps several CTAs resident per SM whatever the number of leaves)
  int lociCap = kMaxBatchLoci;
  size_t smemBudget = 64 * 1024;
  if (const char* e = getenv("GPHOCS_EVAL_SMEM_BUDGET")) smemBudget = (size_t)atol(e);
  if (const char* e = getenv("GPHOCS_EVAL_MAX_LOCI")) lociCap = std::max(1, std::min(kMaxBatchLoci, atoi(e)));
  while (lociCap > 1 && evalSmemBytes(n, lociCap) > smemBudget) lociCap--;
  {
    Batch cur{0, 0, 0, 0, -1, 0};
    auto flush = [&]() { if (cur.numLoci > 0) s->batches.push_back(cur); cur = Batch{0, 0, 0, 0, -1, 0}; };
    for (int l = 0; l < L; l++) {
      const int P = s->colStart[l + 1] - s->colStart[l];
      if (P == 0) { flush(); continue; }  // keeps batches contiguous in locus index
      if (P > kThreads) {
        flush();
        Batch b{l, 1, s->colStart[l], P, (int)scratchCols, 0};
        scratchCols += P;
        s->locusBatch[l] = (int)s->batches.size();
        s->batches.push_back(b);
        continue;
      }
      if (cur.numLoci > 0 && (cur.numCols + P > kThreads || cur.numLoci == lociCap)) flush();
      if (cur.numLoci == 0) { cur.firstLocus = l; cur.firstCol = s->colStart[l]; }
      s->locusBatch[l] = (int)s->batches.size();
      cur.numLoci++;
      cur.numCols += P;
    }
    flush();
  }
  s->numBatches = (int)s->batches.size();
  s->anyOversized = scratchCols > 0;
  for (const Batch& bt : s->batches) s->maxBatchLoci = std::max(s->maxBatchLoci, bt.numLoci);

  // ---- device allocations
  StoreDev& d = s->d;
  d.L = L; d.n = n; d.N = N; d.NI = NI; d.W = W; d.Ct = Ct;
  int *dColStart = nullptr, *dPh = nullptr, *dCnt = nullptr;
  unsigned long long* dWords = nullptr;
  bool ok = true;
  const size_t LN = (size_t)L * N;
  ok = ok && devAlloc(&dColStart, L + 1) == 0 && devAlloc(&dWords, words.size()) == 0 && devAlloc(&dPh, phases.size()) == 0 &&
       devAlloc(&dCnt, cnt.size()) == 0 && devAlloc(&d.clv, (size_t)Ct * NI * 8) == 0 && devAlloc(&d.node, LN) == 0 &&
       devAlloc(&d.saved, LN) == 0 && devAlloc(&d.age, LN) == 0 &&
       devAlloc(&d.svAge, LN) == 0 && devAlloc(&d.root, L) == 0 &&
       devAlloc(&d.savedRoot, L) == 0 && devAlloc(&d.rate, L) == 0 && devAlloc(&d.lnL, L) == 0 &&
       devAlloc(&d.savedLnL, L) == 0 && devAlloc(&d.rootScratch, (size_t)scratchCols * 4) == 0 &&
       devAlloc(&d.ctaSum, s->numBatches) == 0 && devAlloc(&s->dMask, L) == 0 &&
       devAlloc(&s->dBatches, s->numBatches) == 0 && devAlloc(&s->dSum, 1) == 0;
  if (!ok) {
    fprintf(stderr, "gphocs_b200: device allocation failed\n");
    delete s;
    return nullptr;
  }
  d.colStart = dColStart; d.leafWords = dWords; d.grpPhases = dPh; d.grpCount = dCnt; d.active = nullptr; d.evalCounters = nullptr;
  s->deviceBytes = (long long)((size_t)Ct * NI * 64 + words.size() * 8 + (size_t)Ct * 8 + LN * (2 * 8 + 16) + (size_t)L * 36);
  cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  s->ownStream = true;
  cudaMemcpy(dColStart, s->colStart.data(), sizeof(int) * (L + 1), cudaMemcpyHostToDevice);
  cudaMemcpy(dWords, words.data(), words.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dPh, phases.data(), phases.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dCnt, cnt.data(), cnt.size() * 4, cudaMemcpyHostToDevice);
  if (s->numBatches) cudaMemcpy(s->dBatches, s->batches.data(), sizeof(Batch) * s->numBatches, cudaMemcpyHostToDevice);
  cudaMemset(d.clv, 0, (size_t)Ct * NI * 64);
  cudaMemset(d.age, 0, LN * 8); cudaMemset(d.svAge, 0, LN * 8);
  cudaMemset(d.root, 0xff, (size_t)L * 4); cudaMemset(d.savedRoot, 0xff, (size_t)L * 4);
  cudaMemset(d.lnL, 0, (size_t)L * 8); cudaMemset(d.savedLnL, 0, (size_t)L * 8);
  cudaMemset(s->dMask, 0, L);
  {
    std::vector<double> ones(L, 1.0);
    cudaMemcpy(d.rate, ones.data(), (size_t)L * 8, cudaMemcpyHostToDevice);
  }
  s->smemBytes = evalSmemBytes(n, s->maxBatchLoci);
  if (s->smemBytes > 220 * 1024) {
    fprintf(stderr, "gphocs_b200: %d leaves need %zu bytes of shared memory per locus batch\n", n, s->smemBytes);
    delete s;
    return nullptr;
  }
  if (raiseDynamicSmem((const void*)k_eval, s->smemBytes) != 0 || cudaDeviceSynchronize() != cudaSuccess) {
    fprintf(stderr, "gphocs_b200: device initialisation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    delete s;
    return nullptr;
  }
  {
    int perSm = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_eval, kThreads, s->smemBytes);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    s->prefetchAhead = std::max(1, perSm) * std::max(1, sms);
    if (const char* e = getenv("GPHOCS_EVAL_PREFETCH")) if (atoi(e) == 0) s->prefetchAhead = 0;
  }
  // ---- host mirror
  s->hNode.assign(LN, NodeRec{-1, -1, -1, 0, 0}); s->hsNode.assign(LN, NodeRec{-1, -1, -1, 0, 0});
  s->hAge.assign(LN, 0.0); s->hsAge.assign(LN, 0.0);
  cudaMemcpy(d.node, s->hNode.data(), LN * sizeof(NodeRec), cudaMemcpyHostToDevice);
  cudaMemcpy(d.saved, s->hsNode.data(), LN * sizeof(NodeRec), cudaMemcpyHostToDevice);
  s->hRoot.assign(L, -1); s->hSavedRoot.assign(L, -1);
  s->hRate.assign(L, 1.0); s->hLnL.assign(L, 0.0); s->hSavedLnL.assign(L, 0.0);
  return s;
}

extern "C" int gphocsStoreDestroy(GphocsStore* s) {
  if (!s) return 0;
  cudaSetDevice(s->device);
  cudaDeviceSynchronize();
  StoreDev& d = s->d;
  void* ptrs[] = {(void*)d.colStart, (void*)d.leafWords, (void*)d.grpPhases, (void*)d.grpCount, d.clv, d.node, d.saved,
                  d.age, d.svAge, d.root, d.savedRoot, d.rate, d.lnL,
                  d.savedLnL, d.rootScratch, d.ctaSum, s->dMask, s->dBatches, s->dSum};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (s->dTopo32) cudaFree(s->dTopo32);
  if (s->dBadTopo) cudaFree(s->dBadTopo);
  if (s->dOpsAsync) cudaFree(s->dOpsAsync);
  if (s->dBadOps) cudaFree(s->dBadOps);
  s->ops.release(); s->seg.release(); s->status.release(); s->ids.release(); s->f64.release(); s->i16.release();
  if (s->ownStream && s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return 0;
}

extern "C" int gphocsStoreSetStream(GphocsStore* s, void* cudaStream) {
  cudaSetDevice(s->device);
  cudaStreamSynchronize(s->stream);
  if (s->ownStream && s->stream) cudaStreamDestroy(s->stream);
  if (cudaStream) { s->stream = (cudaStream_t)cudaStream; s->ownStream = false; }
  else { cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking); s->ownStream = true; }
  return 0;
}
extern "C" int gphocsStoreNumLoci(const GphocsStore* s) { return s->L; }
extern "C" int gphocsStoreNumLeaves(const GphocsStore* s) { return s->n; }
extern "C" long long gphocsStoreNumColumns(const GphocsStore* s) { return s->Ct; }
extern "C" long long gphocsStoreDeviceBytes(const GphocsStore* s) { return s->deviceBytes; }
extern "C" long long gphocsKernelLaunchCount(void) { return g_launches.load(); }
// stream-ordered copy out of the library's device results, so callers can gather them into their own buffers — on
// the device (e.g. the all-reduce payload of SURVEY.md §8e) or in page-locked host memory — without blocking
extern "C" int gphocsCopyDeviceAsync(void* dst, const void* src, long long bytes, void* cudaStream) {
  CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)cudaStream));
  return 0;
}
// host threads used for staging conversions and the host mirror (defaults to OpenMP's choice, which launchers
// such as torchrun pin to 1 through OMP_NUM_THREADS)
extern "C" int gphocsSetHostThreads(int n) {
  return setHostThreads(n);
}
// page-locked host memory for callers that want their input/output arrays to move at full PCIe speed
extern "C" void* gphocsHostAlloc(long long bytes) {
  void* p = nullptr;
  if (bytes <= 0 || cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) {
    fprintf(stderr, "gphocs_b200: cannot allocate %lld bytes of page-locked host memory\n", bytes);
    return nullptr;
  }
  return p;
}
extern "C" int gphocsHostFree(void* p) {
  if (p) CUDA_TRY(cudaFreeHost(p));
  return 0;
}
static int takeBadOps(GphocsStore* s);
extern "C" int gphocsStoreSync(GphocsStore* s) {
  cudaSetDevice(s->device);
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return takeBadOps(s);
}

// ---- genealogies
// The page-locked route of gphocsStoreSetTrees does not touch the host mirror (updating 100k genealogies costs the
// host as long as the PCIe transfer itself): it marks the mirror's topology, ages and roots stale, and whoever
// reads the mirror next — an edit batch, a getter of the scalar API, gphocsStoreGetTrees, the mirror check — brings
// it up to date from the device copy first.  Flag bytes of the mirror are its own and are kept.
static int refreshMirrorLocked(GphocsStore* s) {
  if (s->lnlStale.load(std::memory_order_acquire)) {
    cudaSetDevice(s->device);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaMemcpy(s->hLnL.data(), s->d.lnL, (size_t)s->L * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(s->hSavedLnL.data(), s->d.savedLnL, (size_t)s->L * sizeof(double), cudaMemcpyDeviceToHost));
    s->lnlStale.store(false, std::memory_order_release);
  }
  if (!s->mirrorStale.load(std::memory_order_acquire)) return 0;
  cudaSetDevice(s->device);
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  const size_t LN = (size_t)s->L * s->N;
  std::vector<NodeRec> nd(LN);
  CUDA_TRY(cudaMemcpy(nd.data(), s->d.node, LN * sizeof(NodeRec), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(s->hAge.data(), s->d.age, LN * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(s->hRoot.data(), s->d.root, (size_t)s->L * sizeof(int), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(s->hRate.data(), s->d.rate, (size_t)s->L * sizeof(double), cudaMemcpyDeviceToHost));
  NodeRec* hn = s->hNode.data();
  const NodeRec* dn = nd.data();
  const bool withFlags = s->flagsStale.load(std::memory_order_acquire);
  parallelFor(0, (long long)LN, [&](long long lo_, long long hi_) {
    for (long long i = lo_; i < hi_; i++) {
      NodeRec rec = hn[i];
      rec.father = dn[i].father; rec.left = dn[i].left; rec.right = dn[i].right;
      if (withFlags) rec.flags = s->debugMirror ? dn[i].flags : (uint8_t)((rec.flags & ~F_SAVED) | (dn[i].flags & F_SAVED));
      hn[i] = rec;
    }
  }, 65536);
  if (withFlags) {   // proposals made on the device only may be pending: their saved copies belong to the mirror too
    CUDA_TRY(cudaMemcpy(s->hsNode.data(), s->d.saved, LN * sizeof(NodeRec), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(s->hsAge.data(), s->d.svAge, LN * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(s->hSavedRoot.data(), s->d.savedRoot, (size_t)s->L * sizeof(int), cudaMemcpyDeviceToHost));
    s->flagsStale.store(false, std::memory_order_release);
  }
  s->mirrorStale.store(false, std::memory_order_release);
  return 0;
}
// for readers that do not hold the store's mutex (the scalar API's getters: one relaxed load on the fast path)
static void ensureMirror(GphocsStore* s) {
  if (__builtin_expect(s->mirrorStale.load(std::memory_order_acquire) || s->lnlStale.load(std::memory_order_acquire), 0)) {
    std::lock_guard<std::mutex> lk(s->mu);
    refreshMirrorLocked(s);
  }
}

static int flushPending(GphocsStore* s);
static int setTreesLocked(GphocsStore* s, int nLoci, const int* locusIds, const int* father, const int* left,
                          const int* right, const double* age, const int* root) {
  const int N = s->N;
  cudaSetDevice(s->device);
  if (nLoci <= 0) return 0;
  if (locusIds)
    for (int k = 0; k < nLoci; k++)
      if (locusIds[k] < 0 || locusIds[k] >= s->L) { fprintf(stderr, "gphocs_b200: locus %d out of range\n", locusIds[k]); return -1; }
  if (!locusIds && nLoci > s->L) { fprintf(stderr, "gphocs_b200: %d genealogies for %d loci\n", nLoci, s->L); return -1; }
  const size_t cnt = (size_t)nLoci * N;
  if (s->opsInFlight) {   // an edit batch may still be reading the shared staging buffers
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    s->opsInFlight = false;
  }
  StoreDev& d = s->d;
  // Page-locked caller arrays covering loci 0..nLoci-1: the DMA engine reads them where they are — ages and roots go
  // straight to their final device arrays, the int32 topology to a scratch that k_set_topology32 packs; nothing is
  // staged and the host mirror is brought up to date only when somebody reads it.
  auto pageLocked = [](const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
  };
  if (!locusIds && nLoci >= 1024 && pageLocked(father) && pageLocked(left) && pageLocked(right) && pageLocked(age) && pageLocked(root)) {
    if (cnt * 3 > s->topo32Cap) {
      if (s->dTopo32) cudaFree(s->dTopo32);
      s->dTopo32 = nullptr;
      s->topo32Cap = 0;
      if (devAlloc(&s->dTopo32, cnt * 3)) return -1;
      s->topo32Cap = cnt * 3;
    }
    CUDA_TRY(cudaMemcpyAsync(s->dTopo32, father, sizeof(int) * cnt, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->dTopo32 + cnt, left, sizeof(int) * cnt, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->dTopo32 + 2 * cnt, right, sizeof(int) * cnt, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(d.age, age, sizeof(double) * cnt, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(d.root, root, sizeof(int) * (size_t)nLoci, cudaMemcpyHostToDevice, s->stream));
    if (!s->dBadTopo && devAlloc(&s->dBadTopo, 1)) return -1;
    CUDA_TRY(cudaMemsetAsync(s->dBadTopo, 0, sizeof(int), s->stream));
    k_set_topology32<<<(unsigned)((cnt + 255) / 256), 256, 0, s->stream>>>(d, s->dTopo32, s->dTopo32 + cnt, s->dTopo32 + 2 * cnt, cnt,
                                                                           s->dBadTopo);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    int bad = 0;
    CUDA_TRY(cudaMemcpyAsync(&bad, s->dBadTopo, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    s->mirrorStale.store(true, std::memory_order_release);   // the mirror follows on demand (refreshMirrorLocked)
    CUDA_TRY(cudaStreamSynchronize(s->stream));   // the caller's arrays are free again when the call returns
    if (bad) { fprintf(stderr, "gphocs_b200: %d node records with ids outside the tree\n", bad); return -1; }
    return 0;
  }
  if (s->i16.reserve(cnt * 3) || s->f64.reserve(cnt) || s->ids.reserve(nLoci) || s->seg.reserve(nLoci)) return -1;
  {   // ids are narrowed to 16 bits below: refuse anything outside the tree before the mirror or the device is touched
    std::atomic<int> badIds{0};
    parallelFor(0, (long long)cnt, [&](long long lo_, long long hi_) {
      int b = 0;
      for (long long i = lo_; i < hi_; i++)
        b += father[i] < -1 || father[i] >= N || left[i] < -1 || left[i] >= N || right[i] < -1 || right[i] >= N;
      if (b) badIds.fetch_add(b);
    }, 65536);
    for (int k = 0; k < nLoci; k++) badIds.fetch_add(root[k] < -1 || root[k] >= N);
    if (badIds.load()) { fprintf(stderr, "gphocs_b200: %d node records with ids outside the tree\n", badIds.load()); return -1; }
  }
  // Loci are converted in chunks by all host threads straight into page-locked staging (int16 topology, fp64 ages,
  // roots, ids) and each chunk's copies are enqueued at once, so PCIe transfers overlap the conversion of the next
  // chunk.  The host mirror is updated in the same pass; flag bytes (buffer selectors) are kept on both sides.
  const int numChunks = nLoci >= 8192 ? 8 : 1;
  for (int c = 0; c < numChunks; c++) {
    const int k0 = (int)((long long)nLoci * c / numChunks), k1 = (int)((long long)nLoci * (c + 1) / numChunks);
    parallelFor(k0, k1, [&](long long lo_, long long hi_) {
    for (int k = (int)lo_; k < (int)hi_; k++) {
      const int l = locusIds ? locusIds[k] : k;
      const size_t o = (size_t)l * N, in = (size_t)k * N;
      int16_t* __restrict__ o3 = s->i16.host + in * 3;
      NodeRec* __restrict__ hn = s->hNode.data() + o;
      const int* __restrict__ fa = father + in;
      const int* __restrict__ le = left + in;
      const int* __restrict__ ri = right + in;
      for (int i = 0; i < N; i++) {     // one 8-byte store per mirror record (flag byte kept), three shorts to staging
        NodeRec rec = hn[i];
        rec.father = o3[3 * i] = (int16_t)fa[i];
        rec.left = o3[3 * i + 1] = (int16_t)le[i];
        rec.right = o3[3 * i + 2] = (int16_t)ri[i];
        hn[i] = rec;
      }
      memcpy(s->f64.host + in, age + in, sizeof(double) * (size_t)N);
      memcpy(s->hAge.data() + o, age + in, sizeof(double) * (size_t)N);
      s->hRoot[l] = s->seg.host[k] = root[k];
      s->ids.host[k] = l;
    }
    }, 256);
    const size_t n0 = (size_t)k0 * N, nn = (size_t)(k1 - k0) * N;
    CUDA_TRY(cudaMemcpyAsync(s->ids.dev + k0, s->ids.host + k0, sizeof(int) * (k1 - k0), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->seg.dev + k0, s->seg.host + k0, sizeof(int) * (k1 - k0), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->i16.dev + n0 * 3, s->i16.host + n0 * 3, sizeof(int16_t) * nn * 3, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->f64.dev + n0, s->f64.host + n0, sizeof(double) * nn, cudaMemcpyHostToDevice, s->stream));
  }
  k_set_trees<<<(unsigned)((cnt + 255) / 256), 256, 0, s->stream>>>(d, s->ids.dev, s->i16.dev, s->f64.dev, s->seg.dev, nLoci);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(s->stream));  // the staging buffers are reused by the next call
  return 0;
}

extern "C" int gphocsStoreSetTrees(GphocsStore* s, int nLoci, const int* locusIds, const int* father, const int* left,
                                   const int* right, const double* age, const int* root) {
  std::lock_guard<std::mutex> lk(s->mu);
  if (flushPending(s)) return -1;   // edits queued by the scalar API belong to the genealogies they were made on
  return setTreesLocked(s, nLoci, locusIds, father, left, right, age, root);
}

// genealogies of loci 0..nLoci-1 in the wire format of the device (16-bit topology triples): three copies and one
// pack kernel; the mirror follows on demand like on the page-locked route above
extern "C" int gphocsStoreSetTreesPacked(GphocsStore* s, int nLoci, const short* topo, const double* age, const int* root) {
  std::lock_guard<std::mutex> lk(s->mu);
  if (flushPending(s)) return -1;
  const int N = s->N;
  cudaSetDevice(s->device);
  if (nLoci <= 0) return 0;
  if (nLoci > s->L) { fprintf(stderr, "gphocs_b200: %d genealogies for %d loci\n", nLoci, s->L); return -1; }
  const size_t cnt = (size_t)nLoci * N;
  if (s->opsInFlight) {
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    s->opsInFlight = false;
  }
  // scratch: the int32 scratch of the page-locked route holds the 16-bit triples as well (3 shorts < 3 ints per node)
  if (cnt * 3 > s->topo32Cap) {
    if (s->dTopo32) cudaFree(s->dTopo32);
    s->dTopo32 = nullptr;
    s->topo32Cap = 0;
    if (devAlloc(&s->dTopo32, cnt * 3)) return -1;
    s->topo32Cap = cnt * 3;
  }
  if (!s->dBadTopo && devAlloc(&s->dBadTopo, 1)) return -1;
  int16_t* dTopo = reinterpret_cast<int16_t*>(s->dTopo32);
  StoreDev& d = s->d;
  CUDA_TRY(cudaMemsetAsync(s->dBadTopo, 0, sizeof(int), s->stream));
  CUDA_TRY(cudaMemcpyAsync(dTopo, topo, sizeof(int16_t) * 3 * cnt, cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaMemcpyAsync(d.age, age, sizeof(double) * cnt, cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaMemcpyAsync(d.root, root, sizeof(int) * (size_t)nLoci, cudaMemcpyHostToDevice, s->stream));
  k_set_topology16<<<(unsigned)((cnt + 255) / 256), 256, 0, s->stream>>>(d, dTopo, cnt, s->dBadTopo);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  int bad = 0;
  CUDA_TRY(cudaMemcpyAsync(&bad, s->dBadTopo, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  s->mirrorStale.store(true, std::memory_order_release);
  CUDA_TRY(cudaStreamSynchronize(s->stream));   // the caller's arrays are free again when the call returns
  if (bad) { fprintf(stderr, "gphocs_b200: %d node records with ids outside the tree\n", bad); return -1; }
  return 0;
}

extern "C" int gphocsStoreGetTrees(GphocsStore* s, int nLoci, const int* locusIds, int* father, int* left, int* right,
                                   double* age, int* root) {
  std::lock_guard<std::mutex> lk(s->mu);
  if (refreshMirrorLocked(s)) return -1;
  const int N = s->N;
  for (int k = 0; k < nLoci; k++) {
    const int l = locusIds ? locusIds[k] : k;
    if (l < 0 || l >= s->L) return -1;
    const size_t o = (size_t)l * N, out = (size_t)k * N;
    for (int i = 0; i < N; i++) {
      father[out + i] = s->hNode[o + i].father;
      left[out + i] = s->hNode[o + i].left;
      right[out + i] = s->hNode[o + i].right;
      age[out + i] = s->hAge[o + i];
    }
    root[k] = s->hRoot[l];
  }
  return 0;
}

// rates of the host mirror (getLocusMutationRate, LocusDataLikelihood.c:382, for many loci)
extern "C" int gphocsStoreGetRates(GphocsStore* s, int nLoci, const int* locusIds, double* rates) {
  std::lock_guard<std::mutex> lk(s->mu);
  for (int k = 0; k < nLoci; k++) {
    const int l = locusIds ? locusIds[k] : k;
    if (l < 0 || l >= s->L) return -1;
    rates[k] = s->hRate[l];
  }
  return 0;
}

// ---- edits
// Ships edit records to the device (grouped by locus, call order kept within a locus) and, while the copy and
// the kernel run, applies the same records to the host mirror with all host threads (loci are independent).
// `mirror` = false when the caller has already applied them to the mirror (scalar API).
static int launchOps(GphocsStore* s, const Op* ops, int nOps, int* outStatus, bool mirror) {
  if (nOps <= 0) return 0;
  cudaSetDevice(s->device);
  if (mirror && refreshMirrorLocked(s)) return -1;   // before anything of this batch reaches the device
  if (s->opsInFlight) {   // the previous batch may still be reading the staging buffers
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    s->opsInFlight = false;
  }
  if (s->ops.reserve(nOps) || s->seg.reserve(nOps + 1) || s->status.reserve(nOps)) return -1;
  bool sorted = true;
  for (int i = 1; i < nOps; i++)
    if (ops[i].locus < ops[i - 1].locus) { sorted = false; break; }
  std::vector<int> order;
  if (!sorted) {  // stable grouping by locus so one device thread replays a locus' records in call order
    order.resize(nOps);
    for (int i = 0; i < nOps; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ops[a].locus < ops[b].locus; });
    parallelFor(0, nOps, [&](long long lo_, long long hi_) {
    for (int i = (int)lo_; i < (int)hi_; i++) s->ops.host[i] = ops[order[i]];
    }, 32768);
  } else {
    parallelFor(0, nOps, [&](long long lo_, long long hi_) {
    for (int i = (int)lo_; i < (int)hi_; i++) s->ops.host[i] = ops[i];
    }, 32768);
  }
  int nSegs = 0;
  for (int i = 0; i < nOps; i++)
    if (i == 0 || s->ops.host[i].locus != s->ops.host[i - 1].locus) s->seg.host[nSegs++] = i;
  s->seg.host[nSegs] = nOps;
  CUDA_TRY(cudaMemcpyAsync(s->ops.dev, s->ops.host, sizeof(Op) * nOps, cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaMemcpyAsync(s->seg.dev, s->seg.host, sizeof(int) * (nSegs + 1), cudaMemcpyHostToDevice, s->stream));
  k_apply_ops<<<(nSegs + 127) / 128, 128, 0, s->stream>>>(s->d, s->ops.dev, s->seg.dev, nSegs, s->status.dev);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  const bool wantDeviceStatus = outStatus && (s->debugMirror || !mirror);
  if (wantDeviceStatus)
    CUDA_TRY(cudaMemcpyAsync(s->status.host, s->status.dev, sizeof(int) * nOps, cudaMemcpyDeviceToHost, s->stream));
  if (mirror) {
    // host mirror (getters must see the proposal when this call returns); statuses come from the same code
    const Op* hops = s->ops.host;
    const int* seg = s->seg.host;
    parallelFor(0, nSegs, [&](long long lo_, long long hi_) {
    for (int g = (int)lo_; g < (int)hi_; g++) {
      const TreeView t = s->hostView(hops[seg[g]].locus);
      for (int o = seg[g]; o < seg[g + 1]; o++) {
        const int st = applyOp(t, hops[o]);
        if (outStatus) outStatus[sorted ? o : order[o]] = st;
      }
    }
    });
  }
  if (mirror || wantDeviceStatus) {
    CUDA_TRY(cudaStreamSynchronize(s->stream));  // staging buffers are reused by the next call
  } else {
    s->opsInFlight = true;   // scalar / fiber path: the evaluation that follows synchronises the stream
  }
  if (wantDeviceStatus) {
    for (int i = 0; i < nOps; i++) {
      int& dst = outStatus[sorted ? i : order[i]];
      if (mirror && dst != s->status.host[i]) {
        fprintf(stderr, "gphocs_b200: host mirror and device disagree on the status of edit %d\n", i);
        return -1;
      }
      dst = s->status.host[i];
    }
  }
  return 0;
}

static int flushPending(GphocsStore* s) {
  if (s->pending.empty()) return 0;
  std::vector<Op> ops;
  ops.swap(s->pending);
  return launchOps(s, ops.data(), (int)ops.size(), nullptr, false);
}

extern "C" int gphocsStoreApplyOps(GphocsStore* s, int nOps, const GphocsOp* ops_, int* outStatus) {
  std::lock_guard<std::mutex> lk(s->mu);
  const Op* ops = reinterpret_cast<const Op*>(ops_);
  if (flushPending(s)) return -1;
  for (int i = 0; i < nOps; i++) {
    const Op& o = ops[i];
    if (o.locus < 0 || o.locus >= s->L) { fprintf(stderr, "gphocs_b200: locus %d out of range\n", o.locus); return -1; }
    bool ok = o.type >= OP_ADJUST_AGE && o.type <= OP_SET_RATE;
    if (o.type == OP_ADJUST_AGE) ok = o.a >= 0 && o.a < s->N;
    if (o.type == OP_SPR) ok = o.a >= 0 && o.a < s->N && o.b >= 0 && o.b < s->N && o.a != o.b;
    if (!ok) { fprintf(stderr, "gphocs_b200: edit %d (type %d, nodes %d, %d) is outside the tree of locus %d\n", i, o.type, o.a, o.b, o.locus); return -1; }
  }
  return launchOps(s, ops, nOps, outStatus, true);
}

// Edit records for the device copy only, without waiting: the records must be sorted by locus (all records of a locus
// adjacent, in call order) and should lie in page-locked memory (gphocsHostAlloc) — the DMA engine reads them where they
// are and they must stay untouched until the stream has passed the copy (the next synchronising call of this store).
// The host mirror behind the scalar API's getters is brought up to date when somebody reads it.  Records outside the
// store or a tree are refused on the device; the next synchronising call (gphocsStoreSync, gphocsStoreEvaluate, ...) reports it.
extern "C" int gphocsStoreApplyOpsAsync(GphocsStore* s, int nOps, const GphocsOp* ops_) {
  std::lock_guard<std::mutex> lk(s->mu);
  if (flushPending(s)) return -1;
  if (nOps <= 0) return 0;
  cudaSetDevice(s->device);
  if ((size_t)nOps > s->opsAsyncCap) {
    if (s->dOpsAsync) { CUDA_TRY(cudaStreamSynchronize(s->stream)); cudaFree(s->dOpsAsync); }
    s->dOpsAsync = nullptr; s->opsAsyncCap = 0;
    if (devAlloc(&s->dOpsAsync, (size_t)nOps + (size_t)nOps / 4)) return -1;
    s->opsAsyncCap = (size_t)nOps + (size_t)nOps / 4;
  }
  if (!s->dBadOps) {
    if (devAlloc(&s->dBadOps, 1)) return -1;
    CUDA_TRY(cudaMemset(s->dBadOps, 0, sizeof(int)));
  }
  CUDA_TRY(cudaMemcpyAsync(s->dOpsAsync, ops_, sizeof(Op) * (size_t)nOps, cudaMemcpyHostToDevice, s->stream));
  k_apply_ops_sorted<<<(nOps + 127) / 128, 128, 0, s->stream>>>(s->d, s->dOpsAsync, nOps, s->dBadOps);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  s->mirrorStale.store(true, std::memory_order_release);
  s->flagsStale.store(true, std::memory_order_release);
  s->lnlStale.store(true, std::memory_order_release);   // OP_REVERT restores lnL on the device
  return 0;
}
// records refused by the device since the last call (0 = none); synchronises the stream
static int takeBadOps(GphocsStore* s) {
  if (!s->dBadOps) return 0;
  int bad = 0;
  CUDA_TRY(cudaMemcpyAsync(&bad, s->dBadOps, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (bad) {
    CUDA_TRY(cudaMemset(s->dBadOps, 0, sizeof(int)));
    fprintf(stderr, "gphocs_b200: %d edit records of gphocsStoreApplyOpsAsync were outside the store, a tree, or out of order\n", bad);
    return -1;
  }
  return 0;
}

extern "C" int gphocsStoreSetRates(GphocsStore* s, int nLoci, const int* locusIds, const double* rates) {
  std::vector<GphocsOp> ops(nLoci);
  for (int k = 0; k < nLoci; k++) ops[k] = GphocsOp{locusIds ? locusIds[k] : k, GPHOCS_OP_SET_RATE, 0, 0, rates[k]};
  return gphocsStoreApplyOps(s, nLoci, ops.data(), nullptr);
}

// ---- evaluation
// masked: evaluate only loci whose mask byte is set; [bLo, bHi] restricts the launch to that range of CTA batches
static int launchEval(GphocsStore* s, int useOld, int onlyLocus, bool masked, int bLo = 0, int bHi = -1, cudaStream_t onStream = nullptr) {
  cudaSetDevice(s->device);
  StoreDev d = s->d;
  d.active = masked ? s->dMask : nullptr;
  if (bHi >= bLo && onlyLocus < 0) {
    k_eval<<<bHi - bLo + 1, kThreads, s->smemBytes, onStream ? onStream : s->stream>>>(d, s->dBatches, bLo, useOld, -1, s->maxBatchLoci, useOld ? s->prefetchAhead : 0);
    g_launches++;
  } else if (onlyLocus >= 0) {
    const int b = s->locusBatch[onlyLocus];
    if (b < 0) return 0;
    k_eval<<<1, kThreads, s->smemBytes, s->stream>>>(d, s->dBatches, b, useOld, onlyLocus, s->maxBatchLoci, 0);
    g_launches++;
  } else if (s->numBatches > 0) {
    k_eval<<<s->numBatches, kThreads, s->smemBytes, s->stream>>>(d, s->dBatches, 0, useOld, -1, s->maxBatchLoci, useOld ? s->prefetchAhead : 0);
    g_launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int gphocsStoreEvaluateDevice(GphocsStore* s, int useOld, void** devLnL, void** devSum) {
  std::lock_guard<std::mutex> lk(s->mu);
  if (flushPending(s)) return -1;
  if (launchEval(s, useOld, -1, false)) return -1;
  s->lnlStale.store(true, std::memory_order_release);   // the mirror's lnL / savedLnL follow on demand
  k_reduce_sum<<<1, 1024, 0, s->stream>>>(s->d.ctaSum, s->numBatches, s->dSum);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  if (devLnL) *devLnL = s->d.lnL;
  if (devSum) *devSum = s->dSum;
  return 0;
}

static int readLnL(GphocsStore* s, int nLoci, const int* locusIds, double* outLnL, bool refreshMirror) {
  // device -> pinned staging -> caller; keeps the host mirror of lnL current
  const int L = s->L;
  if (s->f64.reserve(L + 1)) return -1;
  if (!locusIds) {
    CUDA_TRY(cudaMemcpyAsync(s->f64.host, s->d.lnL, sizeof(double) * L, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (refreshMirror) memcpy(s->hLnL.data(), s->f64.host, sizeof(double) * L);
    if (outLnL) memcpy(outLnL, s->f64.host, sizeof(double) * nLoci);
  } else {
    if (s->ids.reserve(nLoci)) return -1;
    memcpy(s->ids.host, locusIds, sizeof(int) * nLoci);
    CUDA_TRY(cudaMemcpyAsync(s->ids.dev, s->ids.host, sizeof(int) * nLoci, cudaMemcpyHostToDevice, s->stream));
    k_gather_f64<<<(nLoci + 255) / 256, 256, 0, s->stream>>>(s->d.lnL, s->ids.dev, nLoci, s->f64.dev);
    g_launches++;
    CUDA_TRY(cudaMemcpyAsync(s->f64.host, s->f64.dev, sizeof(double) * nLoci, cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    for (int k = 0; k < nLoci; k++) {
      if (refreshMirror) s->hLnL[locusIds[k]] = s->f64.host[k];
      if (outLnL) outLnL[k] = s->f64.host[k];
    }
  }
  return 0;
}

// Host replay of the buffer flips an evaluation performs on the device (debug / mirror-check mode only):
// exactly the dirty nodes and their ancestors, or every internal node for a full evaluation.
static void replayFlipsOnMirror(GphocsStore* s, int l, int useOld) {
  const int n = s->n, N = s->N;
  if (s->colStart[l + 1] == s->colStart[l] || s->hRoot[l] < 0) return;
  TreeView t = s->hostView(l);
  if (!useOld) {
    for (int i = n; i < N; i++) flipClv(t, i);
    return;
  }
  for (int i = 0; i < N; i++) {
    if (!(t.node[i].flags & F_RECALC) || (t.node[i].flags & 0x80)) continue;
    int u = i < n ? t.node[i].father : i;
    while (u >= 0 && !(t.node[u].flags & 0x80)) {  // 0x80: visited in this replay
      flipClv(t, u);
      t.node[u].flags |= 0x80;
      u = t.node[u].father;
    }
  }
  for (int i = 0; i < N; i++) t.node[i].flags &= 0x7f;
}

static int evaluateLocked(GphocsStore* s, int nLoci, const int* locusIds, int useOld, double* outLnL, double* outSum);

extern "C" int gphocsStoreEvaluate(GphocsStore* s, int nLoci, const int* locusIds, int useOld, double* outLnL,
                                   double* outSum) {
  std::lock_guard<std::mutex> lk(s->mu);
  return evaluateLocked(s, nLoci, locusIds, useOld, outLnL, outSum);
}

static int evaluateLocked(GphocsStore* s, int nLoci, const int* locusIds, int useOld, double* outLnL, double* outSum) {
  if (flushPending(s)) return -1;
  cudaSetDevice(s->device);
  const bool all = (locusIds == nullptr);
  if (all && nLoci != s->L) { fprintf(stderr, "gphocs_b200: locusIds == NULL requires nLoci == numLoci\n"); return -1; }
  if (s->f64.reserve((size_t)s->L + 1)) return -1;
  if (s->lnlStale.load(std::memory_order_acquire) && refreshMirrorLocked(s)) return -1;
  // mirror: savedLnL <- lnL for every evaluated locus (.c:440; loci without patterns or tree hold 0 in both)
  if (all) {
    memcpy(s->hSavedLnL.data(), s->hLnL.data(), sizeof(double) * s->L);
  } else {
    for (int k = 0; k < nLoci; k++) {
      const int l = locusIds[k];
      if (l < 0 || l >= s->L) { fprintf(stderr, "gphocs_b200: locus %d out of range\n", l); return -1; }
      s->hSavedLnL[l] = s->hLnL[l];
    }
  }
  if (!all && nLoci == 1) {
    if (launchEval(s, useOld, locusIds[0], false)) return -1;
  } else if (all) {
    if (launchEval(s, useOld, -1, false)) return -1;
  } else {
    if (s->ids.reserve(nLoci)) return -1;
    memcpy(s->ids.host, locusIds, sizeof(int) * nLoci);
    // only the CTA batches that cover the listed loci are launched; the mask keeps their other loci untouched
    int lMin = s->L, lMax = -1, bLo = s->numBatches, bHi = -1;
    for (int k = 0; k < nLoci; k++) {
      const int l = locusIds[k], b = s->locusBatch[l];
      lMin = std::min(lMin, l); lMax = std::max(lMax, l);
      if (b >= 0) { bLo = std::min(bLo, b); bHi = std::max(bHi, b); }
    }
    if (bHi >= bLo) {
      const int mLo = s->batches[bLo].firstLocus, mHi = s->batches[bHi].firstLocus + s->batches[bHi].numLoci;
      CUDA_TRY(cudaMemsetAsync(s->dMask + mLo, 0, mHi - mLo, s->stream));
      CUDA_TRY(cudaMemcpyAsync(s->ids.dev, s->ids.host, sizeof(int) * nLoci, cudaMemcpyHostToDevice, s->stream));
      k_set_mask<<<(nLoci + 255) / 256, 256, 0, s->stream>>>(s->dMask, s->ids.dev, nLoci);
      g_launches++;
      if (launchEval(s, useOld, -1, true, bLo, bHi)) return -1;
    }
  }
  if (outSum && all) {
    k_reduce_sum<<<1, 1024, 0, s->stream>>>(s->d.ctaSum, s->numBatches, s->dSum);
    g_launches++;
    CUDA_TRY(cudaMemcpyAsync(s->f64.host + s->L, s->dSum, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  }
  if (s->debugMirror && refreshMirrorLocked(s)) return -1;
  if (s->debugMirror)
    for (int k = 0; k < nLoci; k++) replayFlipsOnMirror(s, all ? k : locusIds[k], useOld);
  if (readLnL(s, nLoci, locusIds, outLnL, true)) return -1;  // synchronises the stream
  if (outSum) {
    if (all) {
      *outSum = s->f64.host[s->L];
    } else {
      double v = 0.0;
      for (int k = 0; k < nLoci; k++) v += s->hLnL[locusIds[k]];
      *outSum = v;
    }
  }
  return 0;
}

// Compares the host mirror with the device copy; returns the number of mismatching entries (tree fields,
// roots, SAVED bits; with debug mirroring on also SEL/RECALC bits and lnL/savedLnL), or -1 on error.
extern "C" int gphocsStoreCheckMirror(GphocsStore* s) {
  std::lock_guard<std::mutex> lk(s->mu);
  if (flushPending(s)) return -1;
  if (refreshMirrorLocked(s)) return -1;
  cudaSetDevice(s->device);
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  const size_t LN = (size_t)s->L * s->N;
  std::vector<NodeRec> nd(LN);
  std::vector<double> a(LN), lnl(s->L), sv(s->L), rate(s->L);
  std::vector<int> root(s->L), sroot(s->L);
  CUDA_TRY(cudaMemcpy(nd.data(), s->d.node, LN * sizeof(NodeRec), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(a.data(), s->d.age, LN * 8, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(root.data(), s->d.root, (size_t)s->L * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(sroot.data(), s->d.savedRoot, (size_t)s->L * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(lnl.data(), s->d.lnL, (size_t)s->L * 8, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(sv.data(), s->d.savedLnL, (size_t)s->L * 8, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(rate.data(), s->d.rate, (size_t)s->L * 8, cudaMemcpyDeviceToHost));
  int bad = 0;
  const uint8_t mask = s->debugMirror ? (F_SEL | F_RECALC | F_SAVED) : F_SAVED;
  for (size_t i = 0; i < LN; i++) {
    bad += nd[i].father != s->hNode[i].father;
    bad += nd[i].left != s->hNode[i].left;
    bad += nd[i].right != s->hNode[i].right;
    bad += a[i] != s->hAge[i];
    bad += (nd[i].flags & mask) != (s->hNode[i].flags & mask);
  }
  for (int i = 0; i < s->L; i++) {
    bad += root[i] != s->hRoot[i];
    bad += sroot[i] != s->hSavedRoot[i];
    bad += rate[i] != s->hRate[i];
    if (s->debugMirror) {
      bad += lnl[i] != s->hLnL[i];
      bad += sv[i] != s->hSavedLnL[i];
    }
  }
  return bad;
}

extern "C" int gphocsStoreSetDebug(GphocsStore* s, int on) {
  s->debugMirror = on != 0;
  return 0;
}

extern "C" int gphocsStoreGetLnL(GphocsStore* s, int nLoci, const int* locusIds, double* outLnL) {
  std::lock_guard<std::mutex> lk(s->mu);
  if (flushPending(s)) return -1;
  cudaSetDevice(s->device);
  return readLnL(s, nLoci, locusIds, outLnL, false);
}

extern "C" int gphocsStoreGetClv(GphocsStore* s, int locus, int node, int saved, double* out) {
  std::lock_guard<std::mutex> lk(s->mu);
  if (flushPending(s)) return -1;
  cudaSetDevice(s->device);
  if (locus < 0 || locus >= s->L || node < 0 || node >= s->N) return -1;
  const int P = s->colStart[locus + 1] - s->colStart[locus];
  if (node < s->n) {  // leaves are stored as base masks; expand on request
    std::vector<unsigned long long> w(P);
    CUDA_TRY(cudaMemcpy(w.data(), s->d.leafWords + (size_t)(node >> 4) * s->Ct + s->colStart[locus], sizeof(unsigned long long) * P,
                        cudaMemcpyDeviceToHost));
    for (int p = 0; p < P; p++) {
      const unsigned code = (unsigned)(w[p] >> ((node & 15) * 4)) & 15u;
      for (int b = 0; b < 4; b++) out[p * 4 + b] = (code >> b) & 1u ? 1.0 : 0.0;
    }
    return 0;
  }
  NodeRec rec;
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  CUDA_TRY(cudaMemcpy(&rec, s->d.node + (size_t)locus * s->N + node, sizeof(NodeRec), cudaMemcpyDeviceToHost));
  const int buf = (rec.flags & F_SEL) ^ (saved ? 1 : 0);
  const double* src = s->d.clv + (size_t)s->colStart[locus] * s->NI * 8 + (size_t)((node - s->n) * 2 + buf) * P * 4;
  CUDA_TRY(cudaMemcpy(out, src, sizeof(double) * P * 4, cudaMemcpyDeviceToHost));
  return 0;
}

// ======================================================================================= genealogy likelihood
struct GphocsGenealogy {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  int L = 0, Q = 0, C = 0, B = 0, V = 0;
  long long totalEvents = 0;
  bool evaluatedOnce = false;     // a full evaluation has left the lineage counts at the start of every chain (gphocsGenRecalc)
  int maxTileEvents = 0, numCtas = 0;
  GenParams hp{};
  GenDev d{};
  GenParams* dParams = nullptr;
  int* dEvStart = nullptr;
  uint16_t* dPopStart = nullptr;
  double* dEvTime = nullptr;
  uint16_t* dEvCode = nullptr;
  uint8_t* dLineages = nullptr;
  double* dTotals = nullptr;
  size_t evCap = 0;
  Staging<double> out;
  Staging<int> sEs;
  Staging<uint16_t> sPs, sCode;
  int* dRaw32 = nullptr;          // device scratch for int32 event types / ids / chain offsets copied as they are
  long long* dRawStart = nullptr;
  int* dBadEvents = nullptr;
  size_t raw32Cap = 0;
  int* dRcStatus = nullptr;       // gphocsGenRecalcAsync: per-chain status of the last call, checked by the next synchronising call
  double* dRcDelta = nullptr;
  size_t rcAsyncCap = 0;
  int rcAsyncPairs = 0;
  Staging<int> rcInts;            // gphocsGenRecalc: locus ids, population ids, offsets of the new times, status
  Staging<double> rcTimes, rcDelta;
  std::vector<int> postOrder;
};

static int genPostOrder(const int* son0, const int* son1, int C, int pop, int* out) {
  if (pop < C) { out[0] = pop; return 1; }
  int size = genPostOrder(son0, son1, C, son0[pop], out);
  size += genPostOrder(son0, son1, C, son1[pop], out + size);
  out[size] = pop;
  return size + 1;
}

extern "C" GphocsGenealogy* gphocsGenCreate(int device, int numLoci, int numPops, int numCurPops, int numBands,
                                            const int* popFather, const int* popSon0, const int* popSon1,
                                            const int* samplesPerPop) {
  if (numLoci <= 0 || numPops != 2 * numCurPops - 1 || numPops > kMaxPops || numBands < 0 || numBands > kMaxBands) {
    fprintf(stderr, "gphocs_b200: bad genealogy dimensions (pops %d, current %d, bands %d)\n", numPops, numCurPops, numBands);
    return nullptr;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cudaSetDevice(device) != cudaSuccess) {
    fprintf(stderr, "gphocs_b200: no usable CUDA device — this library has no CPU path\n");
    return nullptr;
  }
  GphocsGenealogy* g = new GphocsGenealogy();
  g->device = device; g->L = numLoci; g->Q = numPops; g->C = numCurPops; g->B = numBands;
  g->V = genTotalsLen(numPops, numBands);
  GenParams& p = g->hp;
  memset(&p, 0, sizeof(p));
  p.Q = numPops; p.C = numCurPops; p.B = numBands; p.rootPop = -1;
  for (int i = 0; i < numPops; i++) {
    p.son0[i] = popSon0[i]; p.son1[i] = popSon1[i];
    p.samplesPerPop[i] = i < numCurPops ? samplesPerPop[i] : 0;
    if (popFather[i] < 0) p.rootPop = i;
    p.theta[i] = 1.0;
  }
  if (p.rootPop < 0 || genPostOrder(p.son0, p.son1, numCurPops, p.rootPop, p.postOrder) != numPops) {
    fprintf(stderr, "gphocs_b200: malformed population tree\n");
    delete g;
    return nullptr;
  }
  g->numCtas = (numLoci + kGenTile - 1) / kGenTile;
  const size_t LQ = (size_t)numLoci * numPops, LB = (size_t)numLoci * std::max(numBands, 1);
  bool ok = devAlloc(&g->dParams, 1) == 0 && devAlloc(&g->dEvStart, numLoci + 1) == 0 &&
            devAlloc(&g->dPopStart, (size_t)numLoci * (numPops + 1)) == 0 && devAlloc(&g->d.lnL, numLoci) == 0 &&
            devAlloc(&g->d.coal, LQ) == 0 && devAlloc(&g->d.numCoals, LQ) == 0 && devAlloc(&g->d.mig, LB) == 0 &&
            devAlloc(&g->d.numMigs, LB) == 0 && devAlloc(&g->d.enter, LQ) == 0 &&
            devAlloc(&g->d.ctaTotals, (size_t)g->numCtas * g->V) == 0 &&
            devAlloc(&g->dTotals, g->V) == 0;
  if (!ok) { delete g; return nullptr; }
  cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
  g->ownStream = true;
  cudaMemcpy(g->dParams, &g->hp, sizeof(GenParams), cudaMemcpyHostToDevice);
  return g;
}

extern "C" int gphocsGenDestroy(GphocsGenealogy* g) {
  if (!g) return 0;
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  void* ptrs[] = {g->dParams, g->dEvStart, g->dPopStart, g->dEvTime, g->dEvCode, g->dLineages, g->dTotals, g->d.lnL,
                  g->d.coal, g->d.numCoals, g->d.mig, g->d.numMigs, g->d.ctaTotals, g->d.enter};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (g->dRaw32) cudaFree(g->dRaw32);
  if (g->dRawStart) cudaFree(g->dRawStart);
  if (g->dBadEvents) cudaFree(g->dBadEvents);
  g->out.release(); g->sEs.release(); g->sPs.release(); g->sCode.release();
  g->rcInts.release(); g->rcTimes.release(); g->rcDelta.release();
  if (g->dRcStatus) cudaFree(g->dRcStatus);
  if (g->dRcDelta) cudaFree(g->dRcDelta);
  if (g->ownStream && g->stream) cudaStreamDestroy(g->stream);
  delete g;
  return 0;
}

extern "C" int gphocsGenSetStream(GphocsGenealogy* g, void* cudaStream) {
  cudaSetDevice(g->device);
  cudaStreamSynchronize(g->stream);
  if (g->ownStream && g->stream) cudaStreamDestroy(g->stream);
  if (cudaStream) { g->stream = (cudaStream_t)cudaStream; g->ownStream = false; }
  else { cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking); g->ownStream = true; }
  return 0;
}

extern "C" int gphocsGenSetParams(GphocsGenealogy* g, const double* theta, const double* migRate) {
  cudaSetDevice(g->device);
  for (int p = 0; p < g->Q; p++) {
    g->hp.theta[p] = theta[p];
    g->hp.log2OverTheta[p] = log(2 / theta[p]);  // same expression as patch.c:2712, evaluated by the host libm
  }
  for (int b = 0; b < g->B; b++) {
    g->hp.migRate[b] = migRate[b];
    g->hp.logMigRate[b] = migRate[b] > 0.0 ? log(migRate[b]) : 0.0;
  }
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  CUDA_TRY(cudaMemcpyAsync(g->dParams, &g->hp, sizeof(GenParams), cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  return 0;
}

// device arrays for E events; the capacity is recorded only once all three allocations have succeeded
static int genReserveEvents(GphocsGenealogy* g, size_t E) {
  if (E <= g->evCap) return 0;
  if (g->dEvTime) cudaFree(g->dEvTime);
  if (g->dEvCode) cudaFree(g->dEvCode);
  if (g->dLineages) cudaFree(g->dLineages);
  g->dEvTime = nullptr; g->dEvCode = nullptr; g->dLineages = nullptr;
  g->evCap = 0;
  const size_t cap = E + E / 8;
  if (devAlloc(&g->dEvTime, cap) || devAlloc(&g->dEvCode, cap) || devAlloc(&g->dLineages, cap)) {
    if (g->dEvTime) cudaFree(g->dEvTime);
    if (g->dEvCode) cudaFree(g->dEvCode);
    if (g->dLineages) cudaFree(g->dLineages);
    g->dEvTime = nullptr; g->dEvCode = nullptr; g->dLineages = nullptr;
    return -1;
  }
  g->evCap = cap;
  return 0;
}

extern "C" int gphocsGenSetEvents(GphocsGenealogy* g, const long long* evStart, const int* popStart, const int* evType,
                                  const int* evId, const double* evTime) {
  cudaSetDevice(g->device);
  const int L = g->L, Q = g->Q;
  const long long E = evStart[L] - evStart[0];
  if (E <= 0 || E >= (1ll << 31)) { fprintf(stderr, "gphocs_b200: bad event count %lld\n", E); return -1; }
  g->totalEvents = 0;   // the object holds no snapshot until this call has succeeded
  g->evaluatedOnce = false;
  if (genReserveEvents(g, (size_t)E)) return -1;
  // Page-locked caller arrays: the DMA engine reads them where they are and k_gen_pack narrows them on the device;
  // the host only finds the largest tile (for the shared-memory size) while the copies are in flight.
  auto pageLocked = [](const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
  };
  if (L >= 1024 && pageLocked(evStart) && pageLocked(popStart) && pageLocked(evType) && pageLocked(evId) && pageLocked(evTime)) {
    const size_t nPs = (size_t)L * (Q + 1), need = 2 * (size_t)E + nPs;
    if (need > g->raw32Cap) {
      if (g->dRaw32) cudaFree(g->dRaw32);
      g->dRaw32 = nullptr;
      g->raw32Cap = 0;
      if (devAlloc(&g->dRaw32, need + need / 8)) return -1;
      g->raw32Cap = need + need / 8;
    }
    if (!g->dRawStart && (devAlloc(&g->dRawStart, (size_t)L + 1) || devAlloc(&g->dBadEvents, 1))) return -1;
    const long long first = evStart[0];
    int* dType = g->dRaw32;
    int* dId = g->dRaw32 + E;
    int* dPs32 = g->dRaw32 + 2 * E;
    CUDA_TRY(cudaMemsetAsync(g->dBadEvents, 0, sizeof(int), g->stream));
    CUDA_TRY(cudaMemcpyAsync(dType, evType + first, sizeof(int) * (size_t)E, cudaMemcpyHostToDevice, g->stream));
    CUDA_TRY(cudaMemcpyAsync(dId, evId + first, sizeof(int) * (size_t)E, cudaMemcpyHostToDevice, g->stream));
    CUDA_TRY(cudaMemcpyAsync(dPs32, popStart, sizeof(int) * nPs, cudaMemcpyHostToDevice, g->stream));
    CUDA_TRY(cudaMemcpyAsync(g->dRawStart, evStart, sizeof(long long) * ((size_t)L + 1), cudaMemcpyHostToDevice, g->stream));
    CUDA_TRY(cudaMemcpyAsync(g->dEvTime, evTime + first, sizeof(double) * (size_t)E, cudaMemcpyHostToDevice, g->stream));
    k_gen_pack<<<1184, 256, 0, g->stream>>>(dType, dId, E, dPs32, (long long)nPs, g->dRawStart, L, g->B, g->dEvCode, g->dPopStart,
                                            g->dEvStart, g->dBadEvents);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    int maxTile = 0;
    for (int l0 = 0; l0 < L; l0 += kGenTile) maxTile = std::max(maxTile, (int)(evStart[std::min(L, l0 + kGenTile)] - evStart[l0]));
    g->maxTileEvents = maxTile;
    int badCount = 0;
    CUDA_TRY(cudaMemcpyAsync(&badCount, g->dBadEvents, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    if (badCount) { fprintf(stderr, "gphocs_b200: malformed event snapshot\n"); return -1; }
    const size_t smemDirect = genSmemBytes(g->Q, g->B, g->maxTileEvents);
    if (smemDirect > 200 * 1024) { fprintf(stderr, "gphocs_b200: event tile too large for shared memory (%zu bytes)\n", smemDirect); return -1; }
    if (raiseDynamicSmem((const void*)k_gen_eval, smemDirect)) return -1;
    g->totalEvents = E;
    return 0;
  }
  if (g->sEs.reserve(L + 1) || g->sPs.reserve((size_t)L * (Q + 1)) || g->sCode.reserve((size_t)E)) return -1;
  int* es = g->sEs.host;            // pinned staging: conversions land where the DMA engine reads them
  uint16_t* ps = g->sPs.host;
  uint16_t* code = g->sCode.host;
  int maxTile = 0;
  bool bad = false;
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  // chunks of loci are converted by all host threads into page-locked staging and their copies enqueued at once,
  // so PCIe transfers overlap the conversion of the next chunk
  const int numChunks = L >= 8192 ? 8 : 1;
  for (int c = 0; c < numChunks && !bad; c++) {
    const int l0 = (int)((long long)L * c / numChunks), l1 = (int)((long long)L * (c + 1) / numChunks);
    parallelFor(l0, l1, [&](long long lo_, long long hi_) {
    for (int l = (int)lo_; l < (int)hi_; l++) {
      es[l] = (int)(evStart[l] - evStart[0]);
      const long long nEv = evStart[l + 1] - evStart[l];
      if (nEv > 65535 || nEv < 0) { bad = true; continue; }
      for (int p = 0; p <= Q; p++) ps[(size_t)l * (Q + 1) + p] = (uint16_t)popStart[(size_t)l * (Q + 1) + p];
      for (long long e = evStart[l]; e < evStart[l + 1]; e++) {
        const int t = evType[e], id = evId[e];
        const bool needsBand = (t == EV_IN_MIG || t == EV_BAND_START || t == EV_BAND_END);
        if (t < 0 || t > EV_DUMMY || (needsBand && (id < 0 || id >= g->B))) bad = true;
        code[e - evStart[0]] = (uint16_t)(t | ((needsBand ? id : 0) << 3));
      }
    }
    });
    if (bad) break;
    const long long e0 = evStart[l0] - evStart[0], e1 = evStart[l1] - evStart[0];
    CUDA_TRY(cudaMemcpyAsync(g->dPopStart + (size_t)l0 * (Q + 1), ps + (size_t)l0 * (Q + 1),
                             sizeof(uint16_t) * (size_t)(l1 - l0) * (Q + 1), cudaMemcpyHostToDevice, g->stream));
    CUDA_TRY(cudaMemcpyAsync(g->dEvCode + e0, code + e0, sizeof(uint16_t) * (size_t)(e1 - e0), cudaMemcpyHostToDevice, g->stream));
    CUDA_TRY(cudaMemcpyAsync(g->dEvTime + e0, evTime + evStart[0] + e0, sizeof(double) * (size_t)(e1 - e0), cudaMemcpyHostToDevice, g->stream));
  }
  es[L] = (int)E;
  if (bad) { cudaStreamSynchronize(g->stream); fprintf(stderr, "gphocs_b200: malformed event snapshot\n"); return -1; }
  for (int l0 = 0; l0 < L; l0 += kGenTile) maxTile = std::max(maxTile, es[std::min(L, l0 + kGenTile)] - es[l0]);
  g->maxTileEvents = maxTile;
  CUDA_TRY(cudaMemcpyAsync(g->dEvStart, es, sizeof(int) * (L + 1), cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  const size_t smem = genSmemBytes(g->Q, g->B, g->maxTileEvents);
  if (smem > 200 * 1024) { fprintf(stderr, "gphocs_b200: event tile too large for shared memory (%zu bytes)\n", smem); return -1; }
  if (raiseDynamicSmem((const void*)k_gen_eval, smem)) return -1;
  g->totalEvents = E;
  return 0;
}

// the snapshot in the device's own format: four copies straight into the final arrays and one checking kernel
extern "C" int gphocsGenSetEventsPacked(GphocsGenealogy* g, const int* evStart, const unsigned short* popStart,
                                        const unsigned short* evCode, const double* evTime) {
  cudaSetDevice(g->device);
  const int L = g->L, Q = g->Q;
  const long long E = evStart[L];
  if (evStart[0] != 0 || E <= 0 || E >= (1ll << 31)) { fprintf(stderr, "gphocs_b200: bad event count %lld\n", E); return -1; }
  g->totalEvents = 0;   // the object holds no snapshot until this call has succeeded
  g->evaluatedOnce = false;
  if (genReserveEvents(g, (size_t)E)) return -1;
  if (!g->dBadEvents && devAlloc(&g->dBadEvents, 1)) return -1;
  const size_t nPs = (size_t)L * (Q + 1);
  CUDA_TRY(cudaMemsetAsync(g->dBadEvents, 0, sizeof(int), g->stream));
  CUDA_TRY(cudaMemcpyAsync(g->dEvStart, evStart, sizeof(int) * ((size_t)L + 1), cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaMemcpyAsync(g->dPopStart, popStart, sizeof(uint16_t) * nPs, cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaMemcpyAsync(g->dEvCode, evCode, sizeof(uint16_t) * (size_t)E, cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaMemcpyAsync(g->dEvTime, evTime, sizeof(double) * (size_t)E, cudaMemcpyHostToDevice, g->stream));
  k_gen_check_packed<<<1184, 256, 0, g->stream>>>(g->dEvCode, E, g->dPopStart, g->dEvStart, L, Q, g->B, g->dBadEvents);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  int maxTile = 0;   // while the copies are in flight
  bool bad = false;
  for (int l0 = 0; l0 < L; l0 += kGenTile) {
    const int n = evStart[std::min(L, l0 + kGenTile)] - evStart[l0];
    if (n < 0) bad = true;
    maxTile = std::max(maxTile, n);
  }
  g->maxTileEvents = maxTile;
  int badCount = 0;
  CUDA_TRY(cudaMemcpyAsync(&badCount, g->dBadEvents, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  if (badCount || bad) { fprintf(stderr, "gphocs_b200: malformed event snapshot\n"); return -1; }
  const size_t smem = genSmemBytes(g->Q, g->B, g->maxTileEvents);
  if (smem > 200 * 1024) { fprintf(stderr, "gphocs_b200: event tile too large for shared memory (%zu bytes)\n", smem); return -1; }
  if (raiseDynamicSmem((const void*)k_gen_eval, smem)) return -1;
  g->totalEvents = E;
  return 0;
}

static int genLaunch(GphocsGenealogy* g, bool wantLineages) {
  cudaSetDevice(g->device);
  if (g->totalEvents <= 0) { fprintf(stderr, "gphocs_b200: gphocsGenSetEvents has not been called\n"); return -1; }
  GenDev d = g->d;
  d.L = g->L; d.Q = g->Q; d.B = g->B;
  d.evStart = g->dEvStart; d.popStart = g->dPopStart; d.evTime = g->dEvTime; d.evCode = g->dEvCode;
  d.evLineages = wantLineages ? g->dLineages : nullptr;
  d.params = g->dParams;
  const size_t smem = genSmemBytes(g->Q, g->B, g->maxTileEvents);
  k_gen_eval<<<g->numCtas, kGenThreads, smem, g->stream>>>(d, g->maxTileEvents, g->hp);
  k_gen_reduce<<<g->V, 256, 0, g->stream>>>(g->d.ctaTotals, g->numCtas, g->V, g->dTotals);
  g_launches += 2;
  CUDA_TRY(cudaGetLastError());
  g->evaluatedOnce = true;
  return 0;
}

extern "C" int gphocsGenEvaluateDevice(GphocsGenealogy* g, void** devLnL, void** devTotals) {
  if (genLaunch(g, false)) return -1;
  if (devLnL) *devLnL = g->d.lnL;
  if (devTotals) *devTotals = g->dTotals;
  return g->V;
}

extern "C" int gphocsGenEvaluate(GphocsGenealogy* g, double* lnL, double* coal, int* numCoals, double* mig, int* numMigs,
                                 double* totalCoal, long long* totalNumCoals, double* totalMig, long long* totalNumMigs,
                                 double* sumLnL) {
  if (genLaunch(g, false)) return -1;
  const int L = g->L, Q = g->Q, B = g->B;
  if (g->out.reserve((size_t)L + g->V)) return -1;
  if (lnL) CUDA_TRY(cudaMemcpyAsync(g->out.host, g->d.lnL, sizeof(double) * L, cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaMemcpyAsync(g->out.host + L, g->dTotals, sizeof(double) * g->V, cudaMemcpyDeviceToHost, g->stream));
  if (coal) CUDA_TRY(cudaMemcpyAsync(coal, g->d.coal, sizeof(double) * L * Q, cudaMemcpyDeviceToHost, g->stream));
  if (numCoals) CUDA_TRY(cudaMemcpyAsync(numCoals, g->d.numCoals, sizeof(int) * L * Q, cudaMemcpyDeviceToHost, g->stream));
  if (mig && B) CUDA_TRY(cudaMemcpyAsync(mig, g->d.mig, sizeof(double) * L * B, cudaMemcpyDeviceToHost, g->stream));
  if (numMigs && B) CUDA_TRY(cudaMemcpyAsync(numMigs, g->d.numMigs, sizeof(int) * L * B, cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  if (lnL) memcpy(lnL, g->out.host, sizeof(double) * L);
  const double* t = g->out.host + L;
  if (sumLnL) *sumLnL = t[0];
  for (int p = 0; p < Q; p++) {
    if (totalCoal) totalCoal[p] = t[1 + p];
    if (totalNumCoals) totalNumCoals[p] = (long long)llround(t[1 + Q + p]);
  }
  for (int b = 0; b < B; b++) {
    if (totalMig) totalMig[b] = t[1 + 2 * Q + b];
    if (totalNumMigs) totalNumMigs[b] = (long long)llround(t[1 + 2 * Q + B + b]);
  }
  return 0;
}

// recalcStats (patch.c:2387-2513) for nPairs (locus, population) chains of the resident snapshot whose events kept
// their number and order but changed their elapsed times (rubberBand, patch.c:596-801): evTime holds the new times of
// the listed chains one after the other, timesStart[nPairs + 1] where each chain's begin.  The chains' statistics are
// recomputed as a full evaluation of the updated snapshot would (bit for bit), stored, and deltaLnL[k] is what
// recalcStats returns.  Needs one full evaluation of the snapshot before (it leaves the lineage counts the chains are
// entered with).  Per-locus log-densities and the totals are those of the last full evaluation until the next one.
extern "C" int gphocsGenRecalc(GphocsGenealogy* g, int nPairs, const int* locus, const int* pop, const int* timesStart,
                               const double* evTime, double* deltaLnL) {
  cudaSetDevice(g->device);
  if (g->totalEvents <= 0) { fprintf(stderr, "gphocs_b200: gphocsGenSetEvents has not been called\n"); return -1; }
  if (!g->evaluatedOnce) { fprintf(stderr, "gphocs_b200: gphocsGenRecalc needs a full evaluation of the snapshot first\n"); return -1; }
  if (nPairs <= 0) return 0;
  const size_t nTimes = (size_t)timesStart[nPairs];
  for (int k = 0; k < nPairs; k++)
    if (locus[k] < 0 || locus[k] >= g->L || pop[k] < 0 || pop[k] >= g->Q || timesStart[k + 1] < timesStart[k]) {
      fprintf(stderr, "gphocs_b200: chain %d (locus %d, population %d) is outside the snapshot\n", k, locus[k], pop[k]);
      return -1;
    }
  if (g->rcInts.reserve(4 * (size_t)nPairs + 1) || g->rcTimes.reserve(nTimes) || g->rcDelta.reserve((size_t)nPairs)) return -1;
  CUDA_TRY(cudaStreamSynchronize(g->stream));   // the staging buffers may still feed the previous call
  int* hi = g->rcInts.host;
  memcpy(hi, locus, sizeof(int) * nPairs);
  memcpy(hi + nPairs, pop, sizeof(int) * nPairs);
  memcpy(hi + 2 * nPairs, timesStart, sizeof(int) * ((size_t)nPairs + 1));
  memcpy(g->rcTimes.host, evTime, sizeof(double) * nTimes);
  CUDA_TRY(cudaMemcpyAsync(g->rcInts.dev, hi, sizeof(int) * (3 * (size_t)nPairs + 1), cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaMemcpyAsync(g->rcTimes.dev, g->rcTimes.host, sizeof(double) * nTimes, cudaMemcpyHostToDevice, g->stream));
  GenDev d = g->d;
  d.L = g->L; d.Q = g->Q; d.B = g->B;
  d.evStart = g->dEvStart; d.popStart = g->dPopStart; d.evTime = g->dEvTime; d.evCode = g->dEvCode;
  d.evLineages = nullptr; d.params = g->dParams;
  int* dStatus = g->rcInts.dev + 3 * (size_t)nPairs + 1;
  k_gen_recalc<<<(nPairs + 127) / 128, 128, 0, g->stream>>>(d, g->dEvTime, nPairs, g->rcInts.dev, g->rcInts.dev + nPairs,
                                                            g->rcInts.dev + 2 * nPairs, g->rcTimes.dev, g->rcDelta.dev, dStatus, g->hp);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(g->rcDelta.host, g->rcDelta.dev, sizeof(double) * nPairs, cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaMemcpyAsync(hi + 3 * (size_t)nPairs + 1, dStatus, sizeof(int) * nPairs, cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  for (int k = 0; k < nPairs; k++)
    if (hi[3 * (size_t)nPairs + 1 + k] != 0) {
      fprintf(stderr, "gphocs_b200: chain %d (locus %d, population %d): %s\n", k, locus[k], pop[k],
              hi[3 * (size_t)nPairs + 1 + k] == 1 ? "the number of events differs from the resident chain's" : "more than 8 overlapping migration bands");
      return -1;
    }
  if (deltaLnL) memcpy(deltaLnL, g->rcDelta.host, sizeof(double) * nPairs);
  return 0;
}

// The same without waiting, for arrays in page-locked memory (gphocsHostAlloc) that stay untouched until the next
// synchronising call on this object: the DMA engine reads them where they are; *devDelta = device array of the nPairs
// return values (stream-ordered: copy it out with gphocsCopyDeviceAsync on the same stream).  Chains refused by the
// device (different number of events, ids outside the snapshot) are reported by the next gphocsGenSync / gphocsGenEvaluate.
extern "C" int gphocsGenRecalcAsync(GphocsGenealogy* g, int nPairs, const int* locus, const int* pop, const int* timesStart,
                                    const double* evTime, void** devDelta) {
  cudaSetDevice(g->device);
  if (g->totalEvents <= 0 || !g->evaluatedOnce) {
    fprintf(stderr, "gphocs_b200: gphocsGenRecalcAsync needs a snapshot and one full evaluation of it\n");
    return -1;
  }
  if (nPairs <= 0) return 0;
  const size_t nTimes = (size_t)timesStart[nPairs];
  if (g->rcInts.reserve(4 * (size_t)nPairs + 1) || g->rcTimes.reserve(nTimes)) return -1;
  if ((size_t)nPairs > g->rcAsyncCap) {
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    if (g->dRcStatus) cudaFree(g->dRcStatus);
    if (g->dRcDelta) cudaFree(g->dRcDelta);
    g->dRcStatus = nullptr; g->dRcDelta = nullptr; g->rcAsyncCap = 0;
    if (devAlloc(&g->dRcStatus, (size_t)nPairs) || devAlloc(&g->dRcDelta, (size_t)nPairs)) return -1;
    g->rcAsyncCap = (size_t)nPairs;
  }
  int* di = g->rcInts.dev;
  CUDA_TRY(cudaMemcpyAsync(di, locus, sizeof(int) * (size_t)nPairs, cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaMemcpyAsync(di + nPairs, pop, sizeof(int) * (size_t)nPairs, cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaMemcpyAsync(di + 2 * (size_t)nPairs, timesStart, sizeof(int) * ((size_t)nPairs + 1), cudaMemcpyHostToDevice, g->stream));
  CUDA_TRY(cudaMemcpyAsync(g->rcTimes.dev, evTime, sizeof(double) * nTimes, cudaMemcpyHostToDevice, g->stream));
  GenDev d = g->d;
  d.L = g->L; d.Q = g->Q; d.B = g->B;
  d.evStart = g->dEvStart; d.popStart = g->dPopStart; d.evTime = g->dEvTime; d.evCode = g->dEvCode;
  d.evLineages = nullptr; d.params = g->dParams;
  k_gen_recalc<<<(nPairs + 127) / 128, 128, 0, g->stream>>>(d, g->dEvTime, nPairs, di, di + nPairs, di + 2 * (size_t)nPairs, g->rcTimes.dev,
                                                            g->dRcDelta, g->dRcStatus, g->hp);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  g->rcAsyncPairs = nPairs;
  if (devDelta) *devDelta = g->dRcDelta;
  return 0;
}
// chains the device refused in the last gphocsGenRecalcAsync (call with the stream synchronised)
static int genTakeRecalcStatus(GphocsGenealogy* g) {
  if (g->rcAsyncPairs <= 0) return 0;
  std::vector<int> st((size_t)g->rcAsyncPairs);
  CUDA_TRY(cudaMemcpy(st.data(), g->dRcStatus, sizeof(int) * st.size(), cudaMemcpyDeviceToHost));
  g->rcAsyncPairs = 0;
  for (size_t k = 0; k < st.size(); k++)
    if (st[k] != 0) {
      fprintf(stderr, "gphocs_b200: gphocsGenRecalcAsync: chain %zu was refused (%s)\n", k,
              st[k] == 1 ? "number of events differs from the resident chain's" : st[k] == 2 ? "more than 8 overlapping migration bands"
                                                                                           : "locus or population outside the snapshot");
      return -1;
    }
  return 0;
}

// per-locus statistics as they are stored on the device (after gphocsGenEvaluate / gphocsGenRecalc), without evaluating
extern "C" int gphocsGenGetStats(GphocsGenealogy* g, double* coal, int* numCoals, double* mig, int* numMigs) {
  cudaSetDevice(g->device);
  const size_t LQ = (size_t)g->L * g->Q, LB = (size_t)g->L * g->B;
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  if (coal) CUDA_TRY(cudaMemcpy(coal, g->d.coal, sizeof(double) * LQ, cudaMemcpyDeviceToHost));
  if (numCoals) CUDA_TRY(cudaMemcpy(numCoals, g->d.numCoals, sizeof(int) * LQ, cudaMemcpyDeviceToHost));
  if (mig && LB) CUDA_TRY(cudaMemcpy(mig, g->d.mig, sizeof(double) * LB, cudaMemcpyDeviceToHost));
  if (numMigs && LB) CUDA_TRY(cudaMemcpy(numMigs, g->d.numMigs, sizeof(int) * LB, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int gphocsGenGetLineages(GphocsGenealogy* g, int* numLineages) {
  if (genLaunch(g, true)) return -1;
  std::vector<uint8_t> tmp(g->totalEvents);
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  CUDA_TRY(cudaMemcpy(tmp.data(), g->dLineages, g->totalEvents, cudaMemcpyDeviceToHost));
  for (long long e = 0; e < g->totalEvents; e++) numLineages[e] = tmp[e];
  return 0;
}

extern "C" int gphocsGenSync(GphocsGenealogy* g) {
  cudaSetDevice(g->device);
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  return genTakeRecalcStatus(g);
}

#include "locus_api.inc"
#include "sampler.inc"
#include "ingest.inc"
