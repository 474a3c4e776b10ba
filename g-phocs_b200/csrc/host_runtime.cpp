// host_runtime.cpp — thread teams, parallelFor and the fiber scheduler behind the OpenMP entry points the
// reference host is compiled against (see host_runtime.h).  Plain C++17 + pthreads; x86-64 SysV context switch.
#include "host_runtime.h"

#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#if !defined(__x86_64__)
#error "host_runtime.cpp: the fiber context switch is written for x86-64"
#endif

namespace gphocs {
namespace {

inline void cpuRelax() { __builtin_ia32_pause(); }

// ------------------------------------------------------------------------------------------------ teams
// A team = the calling thread + (n-1) persistent workers that all run job(w) concurrently (w = 0 is the caller).
class Team {
 public:
  ~Team() { shutdown(); }
  // The generation counter and the number of active workers are published in ONE atomic word (generation << 20 |
  // active): a worker decides from a single load whether a generation concerns it, so a worker that sat out generation
  // G and is preempted can neither run G+1's job under G's number nor run it twice.  job_ is read only by workers that
  // are active in the generation they have just observed, and the caller cannot leave that generation (and overwrite
  // job_) before every one of them has decremented pending_.
  void run(int n, const std::function<void(int)>& job) {
    if (n <= 1) { job(0); return; }
    if (n - 1 > kMaxActive) { fprintf(stderr, "gphocs_b200: a team of %d host threads is not supported\n", n); abort(); }
    grow(n - 1);
    job_ = &job;
    pending_.store(n - 1, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(mu_);
      const uint64_t g = (gen_.load(std::memory_order_relaxed) >> kActiveBits) + 1;
      gen_.store(g << kActiveBits | (uint64_t)(n - 1), std::memory_order_release);
    }
    cv_.notify_all();
    job(0);
    int spins = 0;
    while (pending_.load(std::memory_order_acquire) != 0) {
      if (++spins < 4096) cpuRelax(); else sched_yield();
    }
  }
  int size() const { return (int)threads_.size() + 1; }

 private:
  void grow(int workers) {
    while ((int)threads_.size() < workers) {
      const int w = (int)threads_.size() + 1;
      threads_.emplace_back([this, w] { loop(w); });
    }
  }
  void loop(int w) {
    uint64_t seen = 0;   // generation number last observed
    for (;;) {
      int spins = 0;
      uint64_t g;
      while (((g = gen_.load(std::memory_order_acquire)) >> kActiveBits) == seen && !stop_.load(std::memory_order_relaxed)) {
        if (++spins < 20000) { cpuRelax(); continue; }
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return (gen_.load(std::memory_order_acquire) >> kActiveBits) != seen || stop_.load(); });
      }
      if (stop_.load()) return;
      seen = g >> kActiveBits;
      if ((uint64_t)w <= (g & kActiveMask)) {
        (*job_)(w);
        pending_.fetch_sub(1, std::memory_order_release);
      }
    }
  }
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_.store(true);
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
    threads_.clear();
  }
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::atomic<uint64_t> gen_{0};
  std::atomic<int> pending_{0};
  std::atomic<bool> stop_{false};
  const std::function<void(int)>* job_ = nullptr;
  static constexpr int kActiveBits = 20;
  static constexpr uint64_t kActiveMask = (1ull << kActiveBits) - 1;
  static constexpr int kMaxActive = 4096;
};

int g_threads = 0;  // 0: hardware concurrency
std::mutex g_poolMu[2];
Team* poolTeam(int i) { static Team* t[2] = {new Team(), new Team()}; return t[i]; }   // parallelFor (two concurrent callers)
Team* regionTeam() { static Team* t = new Team(); return t; }    // OpenMP regions / fiber workers

int hwThreads() {
  const unsigned h = std::thread::hardware_concurrency();
  return h ? (int)h : 1;
}

}  // namespace

int hostThreads() { return g_threads > 0 ? g_threads : hwThreads(); }
int setHostThreads(int n) {
  g_threads = n < 1 ? 1 : n;
  return g_threads;
}

void parallelFor(long long begin, long long end, const std::function<void(long long, long long)>& f, long long grain) {
  const long long n = end - begin;
  if (n <= 0) return;
  int parts = (int)std::min<long long>(hostThreads(), (n + grain - 1) / grain);
  if (parts <= 1) { f(begin, end); return; }
  // two teams, so that two host threads (e.g. one feeding the data-likelihood store, one the genealogy snapshot)
  // can both run their conversions in parallel; a third concurrent caller runs inline
  for (int i = 0; i < 2; i++) {
    std::unique_lock<std::mutex> lk(g_poolMu[i], std::try_to_lock);
    if (!lk.owns_lock()) continue;
    if (i == 1) parts = std::max(1, parts / 2);
    poolTeam(i)->run(parts, [&](int w) {
      const long long lo = begin + n * w / parts, hi = begin + n * (w + 1) / parts;
      if (lo < hi) f(lo, hi);
    });
    return;
  }
  f(begin, end);
}

// ------------------------------------------------------------------------------------------------ fibers
extern "C" void gphocs_ctx_switch(void** saveSp, void* loadSp);
extern "C" void gphocs_fiber_entry();
asm(R"(
        .text
        .globl  gphocs_ctx_switch
        .type   gphocs_ctx_switch,@function
gphocs_ctx_switch:
        pushq   %rbp
        pushq   %rbx
        pushq   %r12
        pushq   %r13
        pushq   %r14
        pushq   %r15
        movq    %rsp, (%rdi)
        movq    %rsi, %rsp
        popq    %r15
        popq    %r14
        popq    %r13
        popq    %r12
        popq    %rbx
        popq    %rbp
        ret
        .size   gphocs_ctx_switch,.-gphocs_ctx_switch
        .globl  gphocs_fiber_entry
        .type   gphocs_fiber_entry,@function
gphocs_fiber_entry:
        movq    %r12, %rdi
        call    gphocs_fiber_main
        ud2
        .size   gphocs_fiber_entry,.-gphocs_fiber_entry
)");

namespace {

constexpr size_t kStackBytes = 64 * 1024;   // per fiber; the reference's update steps keep little on the stack
constexpr size_t kGuardBytes = 4096;
constexpr int kWave = 16384;                // fibers alive at once

enum FiberState : int { F_READY = 0, F_PARKED = 1, F_DONE = 2 };
struct Fiber {
  void* sp = nullptr;
  int id = 0;
  int state = F_READY;
};

struct Region {
  void (*fn)(void*) = nullptr;
  void* data = nullptr;
  int numFibers = 0;   // what omp_get_num_threads() reports inside the region
  int workers = 1;
};

struct alignas(64) WorkerState {
  void* schedSp = nullptr;
  int parked = 0, live = 0;
};

struct RegionStats {
  long long regions = 0, rounds = 0;
  double regionSeconds = 0.0, flushSeconds = 0.0;
  ~RegionStats() {
    if (regions && getenv("GPHOCS_B200_STATS"))
      fprintf(stderr, "gphocs_b200: %lld parallel regions, %lld scheduling rounds, %.3f s inside regions of which %.3f s in flushes\n",
              regions, rounds, regionSeconds, flushSeconds);
  }
};
RegionStats g_stats;
inline double nowSeconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

FiberHooks g_hooks;
Region g_region;
bool g_inRegion = false;
char* g_stacks = nullptr;
std::vector<WorkerState> g_workers;
std::atomic<int> g_barCount{0};
std::atomic<int> g_barSense{0};
std::atomic<int> g_allDone{0};

thread_local Fiber* t_fiber = nullptr;
thread_local int t_worker = 0;
thread_local int t_tid = 0, t_nth = 1;   // direct (non-fiber) regions

void barrier(int n, int& localSense) {
  localSense ^= 1;
  if (g_barCount.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
    g_barCount.store(0, std::memory_order_relaxed);
    g_barSense.store(localSense, std::memory_order_release);
  } else {
    int spins = 0;
    while (g_barSense.load(std::memory_order_acquire) != localSense) {
      if (++spins < 2048) cpuRelax(); else sched_yield();   // yield: the flushing thread's helpers need the cores
    }
  }
}

void ensureStacks() {
  if (!g_stacks) {
    const size_t total = (size_t)kWave * (kStackBytes + kGuardBytes);
    void* p = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { fprintf(stderr, "gphocs_b200: cannot reserve fiber stacks\n"); exit(-1); }
    g_stacks = (char*)p;
    for (int i = 0; i < kWave; i++) mprotect(g_stacks + (size_t)i * (kStackBytes + kGuardBytes), kGuardBytes, PROT_NONE);
  }
}
char* stackFor(int slot) { return g_stacks + (size_t)slot * (kStackBytes + kGuardBytes) + kGuardBytes; }

void prepare(Fiber& f, int id, int slot) {
  f.id = id;
  f.state = F_READY;
  char* top = stackFor(slot) + kStackBytes;                 // 16-byte aligned
  void** sp = reinterpret_cast<void**>(top) - 7;            // r15 r14 r13 r12 rbx rbp ret
  sp[0] = nullptr; sp[1] = nullptr; sp[2] = nullptr;
  sp[3] = &f;                                               // r12 -> fiber
  sp[4] = nullptr; sp[5] = nullptr;
  sp[6] = reinterpret_cast<void*>(&gphocs_fiber_entry);
  f.sp = sp;
}

// one wave of fibers [w0, w1) on the region's workers
void runWave(int w0, int w1) {
  const int K = g_region.workers;
  std::vector<Fiber> fibers(w1 - w0);
  g_allDone.store(0);
  regionTeam()->run(K, [&](int w) {
    t_worker = w;
    WorkerState& ws = g_workers[w];
    const int n = w1 - w0;
    const int lo = (int)((long long)n * w / K), hi = (int)((long long)n * (w + 1) / K);
    for (int i = lo; i < hi; i++) prepare(fibers[i], w0 + i, i);
    ws.live = hi - lo;
    int sense = g_barSense.load(std::memory_order_acquire);
    for (;;) {
      ws.parked = 0;
      for (int i = lo; i < hi; i++) {
        Fiber& f = fibers[i];
        if (f.state != F_READY) continue;
        t_fiber = &f;
        gphocs_ctx_switch(&ws.schedSp, f.sp);   // runs the fiber until it parks or finishes
        t_fiber = nullptr;
        if (f.state == F_PARKED) ws.parked++;
        else ws.live--;
      }
      barrier(K, sense);
      if (w == 0) {
        int parked = 0, live = 0;
        for (int k = 0; k < K; k++) { parked += g_workers[k].parked; live += g_workers[k].live; }
        // also at the end of the wave: edits queued after the last evaluation must not outlive the region
        g_stats.rounds++;
        if ((parked > 0 || live == 0) && g_hooks.flush) {
          const double t0 = nowSeconds();
          g_hooks.flush();
          g_stats.flushSeconds += nowSeconds() - t0;
        }
        g_allDone.store(live == 0 ? 1 : 0, std::memory_order_release);
      }
      barrier(K, sense);
      if (g_allDone.load(std::memory_order_acquire)) break;
      for (int i = lo; i < hi; i++)
        if (fibers[i].state == F_PARKED) fibers[i].state = F_READY;
    }
  });
}

}  // namespace

extern "C" void gphocs_fiber_main(Fiber* f) {
  g_region.fn(g_region.data);
  f->state = F_DONE;
  WorkerState& ws = g_workers[t_worker];
  gphocs_ctx_switch(&f->sp, ws.schedSp);
  abort();   // a finished fiber is never resumed
}

void setFiberHooks(const FiberHooks& hooks) { g_hooks = hooks; }
bool inFiber() { return t_fiber != nullptr; }
int fiberWorker() { return t_worker; }
int fiberWorkers() { return g_inRegion ? g_region.workers : 1; }
void fiberPark() {
  Fiber* f = t_fiber;
  if (!f) return;
  f->state = F_PARKED;
  WorkerState& ws = g_workers[t_worker];
  gphocs_ctx_switch(&f->sp, ws.schedSp);
}

}  // namespace gphocs

// ------------------------------------------------------------------------------------------------ OpenMP entry points
// What GCC-compiled `#pragma omp parallel for schedule(static)` code needs from its runtime (libgomp ABI).
using namespace gphocs;

extern "C" void GOMP_parallel(void (*fn)(void*), void* data, unsigned numThreads, unsigned /*flags*/) {
  if (g_inRegion) {   // nested region: run it on the calling thread alone
    const int tid = t_tid, nth = t_nth;
    t_tid = 0; t_nth = 1;
    fn(data);
    t_tid = tid; t_nth = nth;
    return;
  }
  int K = numThreads ? (int)numThreads : hostThreads();
  K = std::max(1, std::min(K, hostThreads()));
  const int F = g_hooks.numFibers ? g_hooks.numFibers() : 0;
  g_inRegion = true;
  const double tRegion = nowSeconds();
  g_stats.regions++;
  g_region.fn = fn;
  g_region.data = data;
  g_region.workers = K;
  if ((int)g_workers.size() < K) g_workers.resize(K);
  if (F > 0) {
    // fiber mode: F "threads" of one loop iteration each, in waves
    g_region.numFibers = F;
    ensureStacks();
    for (int w0 = 0; w0 < F; w0 += kWave) runWave(w0, std::min(F, w0 + kWave));
  } else {
    // direct mode: K OS threads, thread ids 0..K-1
    g_region.numFibers = 0;
    regionTeam()->run(K, [&](int w) {
      t_tid = w; t_nth = K;
      fn(data);
      t_tid = 0; t_nth = 1;
    });
  }
  g_stats.regionSeconds += nowSeconds() - tRegion;
  g_inRegion = false;
}

extern "C" int omp_get_thread_num(void) {
  if (Fiber* f = t_fiber) return f->id;
  return t_tid;
}
extern "C" int omp_get_num_threads(void) {
  if (t_fiber) return g_region.numFibers;
  return t_nth;
}
extern "C" int omp_get_max_threads(void) { return hostThreads(); }
extern "C" void omp_set_num_threads(int n) { setHostThreads(n); }
extern "C" void omp_set_dynamic(int) {}
extern "C" int omp_in_parallel(void) { return g_inRegion ? 1 : 0; }

// ------------------------------------------------------------------------------------------------ self-test
// CPU-only check of the scheduler (tests/test_host_runtime.py): a parallel region of `numFibers` one-iteration
// "threads"; every fiber parks `parks` times and adds what the flushes published.  Returns the sum over fibers of
// (id+1) * flushesSeen, or -1 if a fiber observed an inconsistent thread id / count.
namespace {
struct SelfTest {
  int numFibers, parks;
  std::atomic<long long> sum{0};
  std::atomic<int> bad{0};
  std::atomic<int> flushes{0};
};
SelfTest* g_selfTest = nullptr;
void selfTestBody(void* p) {
  SelfTest* st = static_cast<SelfTest*>(p);
  const int nth = omp_get_num_threads(), tid = omp_get_thread_num();
  // what GCC emits for schedule(static): this thread's share of the iteration space [0, numFibers)
  const int q = st->numFibers / nth, r = st->numFibers % nth;
  const int lo = tid < r ? tid * (q + 1) : tid * q + r, hi = lo + (tid < r ? q + 1 : q);
  for (int i = lo; i < hi; i++) {
    volatile char pad[512];   // some live stack across the park
    pad[0] = (char)i;
    long long seen = 0;
    for (int k = 0; k < st->parks; k++) {
      const int before = st->flushes.load();
      fiberPark();
      if (inFiber() && st->flushes.load() <= before) st->bad++;
      seen++;
    }
    if (omp_get_thread_num() != tid || pad[0] != (char)i) st->bad++;
    st->sum += (long long)(i + 1) * seen;
  }
}
}  // namespace

extern "C" long long gphocsFiberSelfTest(int numFibers, int parks, int threads, int useFibers) {
  SelfTest st;
  st.numFibers = numFibers;
  st.parks = parks;
  g_selfTest = &st;
  const FiberHooks saved = g_hooks;
  const int savedThreads = g_threads;
  FiberHooks h;
  h.flush = [&] { st.flushes++; };
  h.numFibers = [&] { return useFibers ? numFibers : 0; };
  setFiberHooks(h);
  setHostThreads(threads);
  GOMP_parallel(selfTestBody, &st, 0, 0);
  setFiberHooks(saved);
  g_threads = savedThreads;
  g_selfTest = nullptr;
  return st.bad.load() ? -1 : st.sum.load();
}
