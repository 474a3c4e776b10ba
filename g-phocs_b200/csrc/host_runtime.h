// host_runtime.h — host-side runtime of libgphocs_b200.so (no CUDA in here):
//
//  * HostPool: a small persistent thread pool with a static-chunk parallelFor.  The library uses it for its own
//    host loops (staging conversions, host mirror) instead of OpenMP, because the library itself provides the
//    OpenMP entry points the reference host is compiled against (below).
//
//  * The fiber scheduler.  The reference host runs its MCMC update steps as `#pragma omp parallel for` loops
//    over loci with schedule(static) (GPhoCS.c:2297,2608,3487,...; MultiCoreUtils.h:8) and calls
//    computeLocusDataLikelihood once per proposal from inside the loop body.  GCC lowers those loops to
//    GOMP_parallel(fn, data, ...) and fn derives its iteration range from omp_get_num_threads() /
//    omp_get_thread_num().  This library exports those entry points: a parallel region over L loci is run as L
//    "threads" of one iteration each, every one a user-level fiber.  A fiber that asks for a likelihood parks;
//    when every fiber of the wave is parked or finished, ONE batched launch evaluates all parked loci and the
//    fibers resume with their results.  The unmodified reference host thereby drives the batched GPU engine —
//    same per-locus RNG streams, same chain — without the loop interchange of INTEGRATION.md §2.
#pragma once
#include <functional>

namespace gphocs {

// ---- thread pool
int hostThreads();
int setHostThreads(int n);
// f(lo, hi) on contiguous chunks of [begin, end), one chunk per host thread; serial when the range is small
void parallelFor(long long begin, long long end, const std::function<void(long long, long long)>& f, long long grain = 2048);

// ---- fibers
struct FiberHooks {
  // called by one thread when every fiber of the wave is parked or done; must serve all parked requests
  std::function<void()> flush;
  // number of fibers (= loci) a parallel region is split into; 0 disables fiber mode
  std::function<int()> numFibers;
};
void setFiberHooks(const FiberHooks& hooks);
bool inFiber();        // is the calling code running on a fiber?
int fiberWorker();     // index of the OS worker thread running the current fiber (0 if none)
int fiberWorkers();    // number of OS worker threads of the current parallel region
void fiberPark();      // park the current fiber until the next flush has completed

}  // namespace gphocs
