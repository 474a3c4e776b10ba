/*
 * gphocs_b200.h — C ABI of libgphocs_b200.so, the B200-native (sm_100a) per-locus likelihood path
 * of G-PhoCS.  Plain pointers and sizes only; no CUDA or torch types cross this boundary
 * (a CUDA stream is passed as void*).
 *
 * Three groups of entry points:
 *
 *  A. The reference's LocusData call surface, same names and signatures as
 *     /root/reference/src/LocusDataLikelihood.h:54-341, so GPhoCS.c / patch.c link against this
 *     library instead of LocusDataLikelihood.o without source changes.
 *     Tree getters are host-memory reads; everything that touches conditional likelihoods runs
 *     on the GPU.  There is no CPU fallback: without a CUDA device these calls fail loudly.
 *
 *  B. The batched engine (GphocsStore): all loci resident in HBM, proposals shipped as 24-byte edit
 *     records, one launch evaluates every locus.  This is what the MCMC update steps of GPhoCS.c
 *     (:2287-4916) call after loop interchange and what bench.py measures.
 *
 *  C. The genealogy likelihood: computeGenetreeStats + recalcStats + gtreeLnLikelihood +
 *     computeTotalStats (/root/reference/src/patch.c:2330,2387,2702,2134) for all loci in one launch
 *     over a flattened snapshot of the host's event chains (patch.h:159-172).
 */
#ifndef GPHOCS_B200_H
#define GPHOCS_B200_H

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ===================================================================================== A. LocusData
 * Replaces src/LocusDataLikelihood.{h,c}.  Line numbers cite LocusDataLikelihood.h. */

typedef struct LOCUS_LIKELIHOOD LocusData; /* opaque, :35 */

/* layout-compatible with src/GenericTree.h:29-39 (only used by copyGenericTreeToLocus) */
#ifndef GPHOCS_B200_NO_GENERIC_TREE
typedef struct GENERIC_BINARY_TREE {
  int numLeaves;
  int rootId;
  char **leafNames;
  int *father;
  int *leftSon;
  int *rightSon;
  double *label1;
  double *label2;
} GenericBinaryTree;
#endif

LocusData *createLocusData(int numLeaves, unsigned short hetMode);                           /* :54  */
int initializeLocusData(LocusData *locusData, char **patternArray, int numPatterns,
                        int *numPhases, int *patternCounts);                                 /* :65  */
int freeLocusData(LocusData *locusData);                                                     /* :74  */
int attachLeaf_UNUSED(LocusData *locusData, int leafId, int target, double age);             /* :87  */
void setLocusMutationRate(LocusData *locusData, double newRate);                             /* :95  */
double getLocusMutationRate(LocusData *locusData);                                           /* :103 */
int computeAllConditionals(LocusData *locusData);                                            /* :114 */
double computeLocusDataLikelihood(LocusData *locusData, unsigned short useOldConditionals);  /* :127 */
double computePatternLogLikelihood(LocusData *locusData, int numPatterns, int *patternIds,
                                   int *patternCounts);                                      /* :137 */
double computeLocusDataLikelihood_deb(LocusData *locusData, unsigned short useOldConditionals); /* :144 */
double addSitePatterns(LocusData *locusData, int numPatterns, int *patternIds, int *patternCounts,
                       unsigned short revertToSaved);                                        /* :155 */
double reduceSitePatterns(LocusData *locusData, int numPatterns, int *patternIds, int *patternCounts,
                          unsigned short revertToSaved);                                     /* :166 */
int checkLocusDataLikelihood(LocusData *locusData);                                          /* :177 */
int revertToSaved(LocusData *locusData);                                                     /* :186 */
int resetSaved(LocusData *locusData);                                                        /* :195 */
int adjustGenNodeAge(LocusData *locusData, int nodeId, double age);                          /* :205 */
double scaleAllNodeAges(LocusData *locusData, double factor);                                /* :215 */
int executeGenSPR(LocusData *locusData, int subtreeRoot, int targetBranch, double age);      /* :228 */
int copyGenericTreeToLocus(LocusData *locusData, GenericBinaryTree *genericTree);            /* :239 */
void printLocusGenTree(LocusData *locusData, FILE *stream, int *nodePops, int *nodeEvents);  /* :250 */
void printLocusDataStats(LocusData *locusData, int maxLogPhases);                            /* :264 */
void printLocusDataPatterns(LocusData *locusData, FILE *outFile);                            /* :276 */
int computePairwiseLCAs(LocusData *locusData, int **lcaMatrix, int *leafArray_aux);          /* :286 */
int getSortedAges(LocusData *locusData, double *ageArray);                                   /* :297 */
double getLocusDataLikelihood(LocusData *locusData);                                         /* :309 */
int getLocusRoot(LocusData *locusData);                                                      /* :317 */
double getNodeAge(LocusData *locusData, int nodeId);                                         /* :325 */
int getNodeFather(LocusData *locusData, int nodeId);                                         /* :333 */
int getNodeSon(LocusData *locusData, int nodeId, unsigned short son);                        /* :341 */

/* Additive: the store that backs every LocusData created so far (built lazily on first use), so a host
 * that has been loop-interchanged can drive the batched engine over the loci it created through A.
 * Locus index = creation order = `gen` of dataState.lociData[gen] (GPhoCS.c:354). */
struct GphocsStore *gpuLociStore(void);
int gpuLocusIndex(LocusData *locusData);

/* ===================================================================================== B. batched engine */

typedef struct GphocsStore GphocsStore;

/* edit record; `type` values below.  One record = one call of the named reference function on `locus`. */
typedef struct GphocsOp {
  int locus;
  int type;
  int a, b;
  double x;
} GphocsOp;
enum {
  GPHOCS_OP_ADJUST_AGE = 0, /* adjustGenNodeAge(locus, a, x)                                   */
  GPHOCS_OP_SPR = 1,        /* executeGenSPR(locus, a, b, x) -> status 0/1/2                    */
  GPHOCS_OP_SCALE_ALL = 2,  /* scaleAllNodeAges(locus, x) without its evaluation               */
  GPHOCS_OP_COMMIT = 3,     /* resetSaved(locus)                                               */
  GPHOCS_OP_REVERT = 4,     /* revertToSaved(locus)                                            */
  GPHOCS_OP_SET_RATE = 5    /* setLocusMutationRate(locus, x)                                  */
};

/* Builds the device-resident store: `numLoci` loci of `numLeaves` leaves; phased patterns in CSR form
 * (pattStart[numLoci+1] rows of `chars`/`numPhases`, unphStart[numLoci+1] rows of `counts`), same
 * meaning as initializeLocusData's arguments.  Returns NULL (and prints why) on error. */
GphocsStore *gphocsStoreCreate(int device, int numLoci, int numLeaves, const long long *pattStart,
                               const long long *unphStart, const char *chars, const int *numPhases,
                               const int *counts);
int gphocsStoreDestroy(GphocsStore *s);
/* all work of the store is enqueued on this CUDA stream (cudaStream_t as void*; NULL = own stream) */
int gphocsStoreSetStream(GphocsStore *s, void *cudaStream);
int gphocsStoreNumLoci(const GphocsStore *s);
int gphocsStoreNumLeaves(const GphocsStore *s);
long long gphocsStoreNumColumns(const GphocsStore *s);
long long gphocsStoreDeviceBytes(const GphocsStore *s);

/* genealogies host -> device.  locusIds NULL = loci 0..nLoci-1.  Arrays are [nLoci][2*numLeaves-1].
 * Page-locked arrays (gphocsHostAlloc) covering all loci are read by the DMA engine where they are.  On that route,
 * and with gphocsStoreSetTreesPacked, the host mirror behind getNodeAge / getNodeFather / getNodeSon / getLocusRoot
 * is not updated by the call: its next reader (those getters, an edit batch, gphocsStoreGetTrees) refreshes it from
 * the device copy, so a loop that only sets genealogies and evaluates never pays for it. */
int gphocsStoreSetTrees(GphocsStore *s, int nLoci, const int *locusIds, const int *father, const int *left,
                        const int *right, const double *age, const int *root);
int gphocsStoreGetTrees(GphocsStore *s, int nLoci, const int *locusIds, int *father, int *left, int *right,
                        double *age, int *root);
/* The same for loci 0..nLoci-1 in the store's own wire format: topo[nLoci][2*numLeaves-1][3] = father, left,
 * right as 16-bit ids (the nodeArray fields of LocusDataLikelihood.c:60-66; -1 where absent).  PCIe is what an
 * end-to-end evaluation waits for, and a host that flattens its genealogies anyway can write 14 instead of 20
 * bytes per node.  Every id must lie in [-1, 2*numLeaves-2] (checked on the device; on -1 the genealogies of the
 * store are undefined until the next successful upload). */
int gphocsStoreSetTreesPacked(GphocsStore *s, int nLoci, const short *topo, const double *age, const int *root);
int gphocsStoreSetRates(GphocsStore *s, int nLoci, const int *locusIds, const double *rates);
int gphocsStoreGetRates(GphocsStore *s, int nLoci, const int *locusIds, double *rates);

/* proposals / accept / reject: applied to the host mirror at once and to the device copy in order.
 * Several records may target one locus (they apply in array order).  outStatus[nOps] may be NULL. */
int gphocsStoreApplyOps(GphocsStore *s, int nOps, const GphocsOp *ops, int *outStatus);
/* the same for the device copy only, without waiting — the delta upload of a host that keeps its own genealogies: records
 * sorted by locus (those of a locus adjacent, in call order), in page-locked memory (gphocsHostAlloc) that stays untouched
 * until the next synchronising call of the store.  The host mirror behind the getters of group A follows on demand.
 * Records outside the store or a tree are refused on the device and reported by gphocsStoreSync. */
int gphocsStoreApplyOpsAsync(GphocsStore *s, int nOps, const GphocsOp *ops);

/* computeLocusDataLikelihood(locus, useOld) for every listed locus (NULL = all) in one launch.
 * outLnL[nLoci] (host) receives the per-locus values, *outSum their sum; either may be NULL. */
int gphocsStoreEvaluate(GphocsStore *s, int nLoci, const int *locusIds, int useOldConditionals,
                        double *outLnL, double *outSum);
/* same, results left on the device: *devLnL = device pointer to lnL[numLoci], *devSum = device pointer
 * to the summed lnL (one double).  For callers that chain a collective on the same stream. */
int gphocsStoreEvaluateDevice(GphocsStore *s, int useOldConditionals, void **devLnL, void **devSum);
int gphocsStoreGetLnL(GphocsStore *s, int nLoci, const int *locusIds, double *outLnL);
/* current (saved=0) / saved (saved=1) conditional likelihoods of one node: out[numPatterns*4] */
int gphocsStoreGetClv(GphocsStore *s, int locus, int node, int saved, double *out);
int gphocsStoreSync(GphocsStore *s);
/* test hooks: with debug on the host mirror also tracks buffer-select / dirty bits and lnL, and
 * gphocsStoreCheckMirror returns the number of entries in which host mirror and device copy differ */
int gphocsStoreSetDebug(GphocsStore *s, int on);
int gphocsStoreCheckMirror(GphocsStore *s);
/* timing hooks for benchmarks: number of kernels this library has launched so far */
long long gphocsKernelLaunchCount(void);
/* number of host threads the library uses for staging conversions and the host mirror (one rank per GPU:
 * give each rank its share of the cores); returns the value in effect */
int gphocsSetHostThreads(int n);
/* CPU-only self-test of the fiber scheduler behind the OpenMP entry points this library exports for the reference
 * host (GOMP_parallel, omp_get_thread_num, omp_get_num_threads, omp_get_max_threads, omp_set_num_threads — the
 * libgomp symbols GPhoCS.o/patch.o reference): sum over fibers of (id+1)*parks, or -1 on inconsistency */
long long gphocsFiberSelfTest(int numFibers, int parks, int threads, int useFibers);
/* page-locked host memory (cudaMallocHost) for callers' input/output arrays, so host<->device copies of the
 * batched calls run at full PCIe speed; plain malloc'd arrays work too, slower */
void *gphocsHostAlloc(long long bytes);
int gphocsHostFree(void *p);
/* stream-ordered copy of device-resident results into a caller's buffer: device memory, or page-locked host memory
 * (gphocsHostAlloc) for a read-back that does not block the host */
int gphocsCopyDeviceAsync(void *dst, const void *src, long long bytes, void *cudaStream);

/* ===================================================================================== C. genealogy likelihood */

typedef struct GphocsGenealogy GphocsGenealogy;

/* event types, numerically equal to the reference's enum event_type (patch.h:159) */
enum { GPHOCS_EV_COAL = 0, GPHOCS_EV_IN_MIG, GPHOCS_EV_OUT_MIG, GPHOCS_EV_MIG_BAND_START,
       GPHOCS_EV_MIG_BAND_END, GPHOCS_EV_SAMPLES_START, GPHOCS_EV_END_CHAIN, GPHOCS_EV_DUMMY };

/* population tree: father/son0/son1 [numPops] (-1 where absent), samplesPerPop[numCurPops] haploid leaves */
GphocsGenealogy *gphocsGenCreate(int device, int numLoci, int numPops, int numCurPops, int numBands,
                                 const int *popFather, const int *popSon0, const int *popSon1,
                                 const int *samplesPerPop);
int gphocsGenDestroy(GphocsGenealogy *g);
int gphocsGenSetStream(GphocsGenealogy *g, void *cudaStream);
/* model parameters (change every UpdateTheta / UpdateMigRates): theta[numPops], migRate[numBands] */
int gphocsGenSetParams(GphocsGenealogy *g, const double *theta, const double *migRate);
/* flattened snapshot of event_chains[gen] for all loci: evStart[numLoci+1]; popStart[numLoci][numPops+1]
 * relative to the locus' first event; per event type, id (migration band for mig/band events, as
 * recalcStats resolves it at patch.c:2425) and elapsed_time. */
int gphocsGenSetEvents(GphocsGenealogy *g, const long long *evStart, const int *popStart, const int *evType,
                       const int *evId, const double *evTime);
/* The same snapshot in the device's own format (10 instead of 16 bytes per event on the wire): evStart[numLoci+1]
 * 32-bit with evStart[0] == 0, popStart 16-bit, evCode = type | band << 3 (band = 0 unless the event is
 * IN_MIG / MIG_BAND_START / MIG_BAND_END), elapsed times as above.  Malformed codes or offsets: -1, and the object
 * holds no snapshot until the next successful call. */
int gphocsGenSetEventsPacked(GphocsGenealogy *g, const int *evStart, const unsigned short *popStart,
                             const unsigned short *evCode, const double *evTime);
/* computeGenetreeStats + gtreeLnLikelihood for every locus, computeTotalStats over them.
 * Host outputs (any may be NULL): lnL[numLoci]; per-locus stats coal[numLoci][numPops],
 * numCoals[numLoci][numPops], mig[numLoci][numBands], numMigs[numLoci][numBands];
 * totals: totalCoal[numPops], totalNumCoals[numPops], totalMig[numBands], totalNumMigs[numBands], *sumLnL */
int gphocsGenEvaluate(GphocsGenealogy *g, double *lnL, double *coal, int *numCoals, double *mig, int *numMigs,
                      double *totalCoal, long long *totalNumCoals, double *totalMig, long long *totalNumMigs,
                      double *sumLnL);
/* device-resident variant: *devTotals points at a packed vector
 * [sumLnL, totalCoal[numPops], totalNumCoals[numPops] (as exact doubles), totalMig[numBands], totalNumMigs[numBands]]
 * = the all-reduce payload of SURVEY.md §8e. Returns its length in doubles. */
int gphocsGenEvaluateDevice(GphocsGenealogy *g, void **devLnL, void **devTotals);
/* recalcStats (patch.c:2387-2513), the incremental half of the path: nPairs (locus, population) chains of the resident
 * snapshot whose events kept their number and order but changed their elapsed times — what rubberBand (patch.c:596-801)
 * does for every UpdateTau proposal.  evTime = the new elapsed times of the listed chains one after the other,
 * timesStart[nPairs + 1] = where each chain's begin.  Statistics of the chains are recomputed bit for bit as a full
 * evaluation of the updated snapshot would and stored; deltaLnL[k] = recalcStats' return value (minus (mig - old) * rate
 * at every MIG_BAND_END in chain order, then minus (coal - old) / theta).  A rejected proposal sends the old times back,
 * as rubberBandRipple(gen, 1) does.  Needs one gphocsGenEvaluate of the snapshot before. */
int gphocsGenRecalc(GphocsGenealogy *g, int nPairs, const int *locus, const int *pop, const int *timesStart,
                    const double *evTime, double *deltaLnL);
/* the same without waiting: arrays in page-locked memory (gphocsHostAlloc), untouched until the next synchronising call;
 * *devDelta = device array of the return values, stream-ordered (gphocsCopyDeviceAsync).  Refused chains (different
 * number of events, ids outside the snapshot) are reported by the next gphocsGenSync. */
int gphocsGenRecalcAsync(GphocsGenealogy *g, int nPairs, const int *locus, const int *pop, const int *timesStart,
                         const double *evTime, void **devDelta);
/* per-locus statistics as stored on the device after gphocsGenEvaluate / gphocsGenRecalc (any pointer may be NULL) */
int gphocsGenGetStats(GphocsGenealogy *g, double *coal, int *numCoals, double *mig, int *numMigs);
/* num_lineages per event as recalcStats leaves it (patch.c:2405); host array [total events] */
int gphocsGenGetLineages(GphocsGenealogy *g, int *numLineages);
int gphocsGenSync(GphocsGenealogy *g);

/* ===================================================================================== D. device-resident MCMC steps
 * the update steps of GPhoCS.c run for all loci per launch with no host round trip inside a sweep —
 * UpdateGB_InternalNode (:2287), UpdateGB_MigSPR (:2598), UpdateTheta (:3035), UpdateTau (:3224), mixing (:4688).
 * Migration bands (gphocsSamplerSetMigration: UpdateGB_MigrationNode :2437, UpdateMigRates :3110), estimated sample
 * ages and locus-rate variation (gphocsSamplerSetAncient: UpdateSampleAge :4006, UpdateLocusRate :4598) are covered.
 * The sampler edits the device copy of the store's genealogies; call gphocsSamplerDownload before reading them
 * through the store or the LocusData getters. */
typedef struct GphocsSampler GphocsSampler;
/* population tree as in gphocsGenCreate; theta[numPops], tau[numPops] (ancestral entries used); Gamma(alpha, beta)
 * priors per population (thetaPrior / agePrior of PopulationTree.h:80-101); nodePop[numLoci][2n-1] = nodePops
 * (patch.h:123).  The store must hold the genealogies (gphocsStoreSetTrees). */
GphocsSampler *gphocsSamplerCreate(GphocsStore *s, int numPops, int numCurPops, const int *popFather, const int *popSon0,
                                   const int *popSon1, const int *samplesPerPop, const double *theta, const double *tau,
                                   const double *thetaAlpha, const double *thetaBeta, const double *tauAlpha,
                                   const double *tauBeta, const int *nodePop, unsigned long long seed);
int gphocsSamplerDestroy(GphocsSampler *sm);
/* loci sharded over ranks (one process per GPU): fn(buf, count, ctx) sums buf over all ranks in place
 * (an NCCL all-reduce of < 1 KB); it is applied to every reduced vector before a global decision, so all ranks keep
 * identical theta / tau.  locusOffset = global index of this rank's first locus (random streams are per global locus). */
int gphocsSamplerSetAllReduce(GphocsSampler *sm, int (*fn)(double *, int, void *), void *ctx, long long locusOffset);
/* Migration bands (MigrationBand, PopulationTree.h:60-70; band = source -> target backwards in time) with their
 * rates and Gamma(alpha, beta) rate priors, and the migration events of every genealogy (genetree_migs,
 * patch.h:138-148): numMigs[numLoci]; migBranch (node below the branch), migBand, migAge are [numLoci][10]
 * (MAX_MIGS).  Switches on UpdateGB_MigrationNode (GPhoCS.c:2437), UpdateMigRates (:3110), migration in the SPR
 * re-simulation (traceLineage, patch.c:886) and band handling in the split-time and mixing steps.  Up to 32
 * bands and 80 leaves.  Trace rows gain the migration rates after the taus. */
int gphocsSamplerSetMigration(GphocsSampler *sm, int numBands, const int *bandSrc, const int *bandTgt, const double *migRate,
                              const double *migAlpha, const double *migBeta, const int *numMigs, const int *migBranch,
                              const int *migBand, const double *migAge);
/* Multi-GPU without a host hook: the library opens its own NCCL communicator (libnccl.so.2, bound at run time) over
 * the ranks that share the model.  One rank obtains a unique id and distributes the 128 bytes; every rank then calls
 * gphocsSamplerInitNccl(sm, id, rank, worldSize, global index of its first locus).  The per-proposal sums
 * (GPhoCS.c:3807-3836, 4796-4803) and the totals (patch.c:2134-2164) are then summed on the device and all-reduced on
 * the sampler's stream; only the reduced vector crosses PCIe. */
int gphocsNcclUniqueId(char *out128);
int gphocsSamplerInitNccl(GphocsSampler *sm, const char *uniqueId128, int rank, int worldSize, long long locusOffset);
/* finetune-mig-time, finetune-mig-rate */
/* new migration rates for all bands, e.g. the host's draw at iteration start-mig (GPhoCS.c:1738-1757) */
int gphocsSamplerSetMigRates(GphocsSampler *sm, const double *migRate);
int gphocsSamplerSetMigFinetunes(GphocsSampler *sm, double migTime, double migRate);
/* Ancient samples and rate variation (BASELINE.json configs[4]).  tau[p < numCurPops] given at creation is the age of
 * population p's samples (pops[p]->sampleAge, PopulationTree.h:97; the leaves' ages in the store must agree).
 * estimate[numCurPops] marks the sample ages that are parameters (`age <x> e`, MCMCcontrol.c:893-910; UpdateSampleAge,
 * GPhoCS.c:4006); finetune[numCurPops] their step sizes (NULL or <= 0: finetune-tau, as MCMCcontrol.c:960).  Their
 * Gamma prior is tauAlpha/tauBeta[p] given at creation (the reference has 0, 0: PopulationTree.c:121).
 * locusRateFinetune > 0 turns on the locus-rate move (UpdateLocusRate, GPhoCS.c:4598) under a Dirichlet(rateAlpha)
 * prior (`locus-mut-rate VAR <alpha>`, finetune-locus-rate).  Trace rows gain the estimated sample ages and the
 * standard deviation of the locus rates before the two likelihood columns. */
int gphocsSamplerSetAncient(GphocsSampler *sm, const int *estimate, const double *finetune, double locusRateFinetune,
                            double rateAlpha);
/* finetune-coal-time, finetune-theta, finetune-tau, finetune-mixing of the control file (MCMCcontrol.c:575-787) */
int gphocsSamplerSetFinetunes(GphocsSampler *sm, double coalTime, double theta, double tau, double mixing);
/* `iterations` MCMC iterations; trace (may be NULL): one row per iteration of gphocsSamplerTraceWidth() doubles =
 * [theta (numPops), tau of ancestral populations, sum of data lnL, sum of genealogy lnL] */
int gphocsSamplerIterate(GphocsSampler *sm, int iterations, double *trace);
int gphocsSamplerTraceWidth(const GphocsSampler *sm);
/* The trace file performMCMC writes (GPhoCS.c:1255-1313 header, 1762-1769 rows; printParamVals :746): once opened,
 * gphocsSamplerIterate appends a row every (sampleSkip+1)-th iteration — iteration, parameters times their print
 * factor (thetaTauPrint = tau-theta-print, migRatePrint = mig-rate-print), mean full log-likelihood per locus, data
 * log-likelihood.  popNames[numPops].  With several GPUs one rank opens the trace. */
int gphocsSamplerOpenTrace(GphocsSampler *sm, const char *path, const char *const *popNames, double thetaTauPrint,
                           double migRatePrint, int sampleSkip);
int gphocsSamplerCloseTrace(GphocsSampler *sm);
/* The coalescence-time and SPR sweeps (UpdateGB_InternalNode GPhoCS.c:2287, UpdateGB_MigSPR :2598) of a model without
 * migration bands run as ONE launch in which a CTA keeps its batch of loci for both sweeps (up to 32 leaves); models
 * with migration bands and larger genealogies take the stepwise route: a proposal launch and an evaluation launch per
 * node.  1 forces the stepwise route everywhere, 0 (default) restores the choice above.  Same random streams and
 * arithmetic: both routes give the same chain bit for bit. */
int gphocsSamplerSetStepwise(GphocsSampler *sm, int on);
/* accepted[10], proposed[10] for {coalescence time, SPR, theta, tau, mixing, migration rate, migration time,
 * (proposed only) split-time moves rejected for a migration conflict, locus rate (pairs of loci), sample age} */
int gphocsSamplerGetState(GphocsSampler *sm, double *theta, double *tau, long long *accepted, long long *proposed);
/* accounting for the roofline of an MCMC iteration: out2[0] = incremental locus evaluations, out2[1] = their algorithmic
 * bytes (32 * P * (2k + 1) with k conditional vectors recomputed) since the previous call with reset != 0;
 * the first call switches the accounting on */
int gphocsSamplerEvalCounters(GphocsSampler *sm, unsigned long long *out2, int reset);
/* checkAll (patch.c:2745) on the device: returns structural violations; largest relative deviation of the
 * incrementally maintained statistics / data log-likelihoods from a recomputation from scratch */
int gphocsSamplerCheck(GphocsSampler *sm, double *maxStatErr, double *maxLnLErr);
/* per-locus statistics (GENETREE_STATS, patch.h:48-51) as the sampler holds them: coal[numLoci][numPops],
 * numCoals[numLoci][numPops], mig[numLoci][numBands], numMigs[numLoci][numBands]; any pointer may be NULL */
int gphocsSamplerGetStats(GphocsSampler *sm, double *coal, int *numCoals, double *mig, int *numMigs);
/* brings the store's host mirror up to date and returns nodePop[numLoci][2n-1] (may be NULL) */
int gphocsSamplerDownload(GphocsSampler *sm, int *nodePop);

/* ===================================================================================== E. alignment ingest
 * The step that produces initializeLocusData's arguments: readSeqFile + processLocusAlignment +
 * cannonizeJCpattern (AlignmentProcessor.c:468-990, 1595-1655) and, per locus, processHetPatterns with
 * computeHetSymmetryBreaks and getAllPhases (:998-1158, 1706-1895, 2242-2290), as called by processAlignments
 * (GPhoCS.c:258-440).  The host parses the text; canonisation, pattern counting in order of first appearance, the
 * greedy choice of arbitrarily phased het genotypes and the phase expansion run on the device for all loci at once.
 * Limits: 64 haploid sample slots, 1024 distinct site patterns per locus, 2^20 phasings per pattern. */
typedef struct GphocsAlignment GphocsAlignment;
/* sampleNames[numSamples] as dataSetup.sampleNames: one entry per haploid slot, "" (or NULL) for the second slot
 * of a diploid sample (MCMCcontrol.c:850-880).  numLociToRead <= 0: all loci of the file.  NULL and a message on
 * stderr where the reference returns -1 (missing file, short/long sequence, illegal base, ambiguity code in a haploid
 * sample, a sample of the control file that never occurs, ...). */
GphocsAlignment *gphocsReadSeqFile(const char *seqFileName, int numSamples, const char *const *sampleNames,
                                   int numLociToRead, int device);
/* processHetPatterns alone, for numLoci loci at once: canonical patterns [start[numLoci]][numSamples] (characters of
 * "TCAGYWKMSRVDBHN"), their counts, isDiploid[numSamples] (AlignmentData.isDiploid) */
GphocsAlignment *gphocsPhasePatterns(int numLoci, int numSamples, const unsigned char *isDiploid, const int *start,
                                     const char *patterns, const int *counts, int breakSymmetries, int device);
int gphocsAlignmentDims(const GphocsAlignment *a, int *numLoci, int *numSamples, int *numPhasedPatterns, int *numPatterns);
/* CSR offsets pattStart/unphStart [numLoci+1]; chars [numPhasedPatterns][numSamples], numPhases [numPhasedPatterns],
 * counts [numPatterns]: initializeLocusData's arguments for every locus, gphocsStoreCreate's layout; canon
 * [numPatterns][numSamples]: the canonical patterns before phasing (AlignmentData.patternArray rows).  Any may be NULL. */
int gphocsAlignmentGet(const GphocsAlignment *a, long long *pattStart, long long *unphStart, char *chars, int *numPhases,
                       int *counts, char *canon);
const char *gphocsAlignmentLocusName(const GphocsAlignment *a, int locus);
/* seconds parsing text / host-to-device / inside the kernels / device-to-host; bytes of symbol rows and of output */
int gphocsAlignmentTimings(const GphocsAlignment *a, double *parse, double *h2d, double *kernel, double *d2h,
                           long long *rawBytes, long long *outBytes);
void gphocsAlignmentFree(GphocsAlignment *a);
/* createLocusData + initializeLocusData for every locus of the alignment (GPhoCS.c:354-403) */
GphocsStore *gphocsStoreFromAlignment(const GphocsAlignment *a, int device);

/* The reference's own entry points for this step (src/AlignmentProcessor.h:137,182,100,124), exported with the
 * same names and meaning so that G-PhoCS links without AlignmentProcessor.o.  They fill and free
 * the host program's `AlignmentData` (struct ALIGNMENT_DATA_STRUCT in AlignmentProcessor.h). */
#ifndef ALIGNMENT_PROCESSOR_H
int readSeqFile(const char *seqFileName, int numSamples, char **sampleNames, int numLociToRead);
int processHetPatterns(char **patternArray, int *patternCounts, int numPatterns, unsigned short breakSymmetries,
                       char ***phasedPatternArray_ptr, int **numPhasesArray_ptr, int *maxNumPhasedPatterns);
int freeAlignmentData(void);
void printAlignmentError(void);
#endif

#ifdef __cplusplus
}
#endif
#endif /* GPHOCS_B200_H */
