/*
 * ref_harness.c — TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * A thin driver that is compiled TOGETHER with the unmodified G-PhoCS reference sources
 * (taken where they lie under /root/reference/src, see oracle/Makefile) into
 * oracle/_ref/libgphocs_ref.so.  It exposes, through a flat C ABI that ctypes can call:
 *   - the reference set-up sequence of main()            (GPhoCS.c:146-237)
 *   - read-only accessors over the reference's global state (patch.h:117-186, GPhoCS.h:47)
 *     so per-locus trees, phased patterns, data lnL, event chains, coal/mig statistics
 *     and genealogy lnL can be dumped as golden vectors
 *   - OpenMP timing loops for the CPU baseline            
 * Nothing here restates reference algorithms: every number comes from the reference's own
 * functions.  The only interposition is initializeLocusData (renamed to
 * ref_initializeLocusData when LocusDataLikelihood.c is compiled) so the phased pattern
 * arrays, which processAlignments frees right after use (GPhoCS.c:424-427), can be recorded.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>
#include <sys/time.h>

#include "utils.h"
#include "MCMCcontrol.h"
#include "AlignmentProcessor.h"
#include "GenericTree.h"
#include "PopulationTree.h"
#include "LocusDataLikelihood.h"
#include "MultiCoreUtils.h"
#include "GPhoCS.h"
#include "patch.h"

/* reference functions without a prototype in the headers we include */
extern int initGeneralInfo();
extern int readControlFile(char *controlFileName);
extern int checkSettings();
extern int finalizeNumParameters();
extern int initializeMCMC();
extern void initRandomGenerator(int nNumLoci, unsigned int unSeed);
extern int ref_initializeLocusData(LocusData *locusData, char **patternArray, int numPatterns,
                                   int *numPhases, int *patternCounts);

/* ------------------------------------------------------------------ pattern recorder */
typedef struct {
  LocusData *locus;
  int numLeaves;
  int numPatterns;  /* phased */
  int numUnphased;
  char *chars;      /* [numPatterns][numLeaves] */
  int *numPhases;   /* [numPatterns] */
  int *counts;      /* [numUnphased] */
} RecordedLocus;

static RecordedLocus *recorded = NULL;
static int numRecorded = 0, capRecorded = 0;
static int recordLeaves = 0;

void refh_set_record_leaves(int n) { recordLeaves = n; }

/* called by GPhoCS.c:403 in place of the reference's initializeLocusData */
int initializeLocusData(LocusData *locusData, char **patternArray, int numPatterns, int *numPhases,
                        int *patternCounts) {
  int n = recordLeaves > 0 ? recordLeaves : dataSetup.numSamples;
  if (numRecorded == capRecorded) {
    capRecorded = capRecorded ? 2 * capRecorded : 1024;
    recorded = (RecordedLocus *)realloc(recorded, capRecorded * sizeof(RecordedLocus));
  }
  RecordedLocus *r = &recorded[numRecorded++];
  r->locus = locusData;
  r->numLeaves = n;
  r->numPatterns = numPatterns;
  r->chars = (char *)malloc((size_t)(numPatterns > 0 ? numPatterns : 1) * n);
  r->numPhases = (int *)malloc(sizeof(int) * (numPatterns > 0 ? numPatterns : 1));
  int u = 0;
  for (int p = 0; p < numPatterns; p++) {
    memcpy(r->chars + (size_t)p * n, patternArray[p], n);
    r->numPhases[p] = numPhases[p];
    if (numPhases[p] > 0) u++;
  }
  r->numUnphased = u;
  r->counts = (int *)malloc(sizeof(int) * (u > 0 ? u : 1));
  for (int i = 0; i < u; i++) r->counts[i] = patternCounts ? patternCounts[i] : 0;
  return ref_initializeLocusData(locusData, patternArray, numPatterns, numPhases, patternCounts);
}

int refh_recorded_count(void) { return numRecorded; }
int refh_recorded_dims(int i, int *numLeaves, int *numPatterns, int *numUnphased) {
  if (i < 0 || i >= numRecorded) return -1;
  *numLeaves = recorded[i].numLeaves;
  *numPatterns = recorded[i].numPatterns;
  *numUnphased = recorded[i].numUnphased;
  return 0;
}
int refh_recorded_get(int i, char *chars, int *numPhases, int *counts) {
  if (i < 0 || i >= numRecorded) return -1;
  RecordedLocus *r = &recorded[i];
  memcpy(chars, r->chars, (size_t)r->numPatterns * r->numLeaves);
  memcpy(numPhases, r->numPhases, sizeof(int) * r->numPatterns);
  memcpy(counts, r->counts, sizeof(int) * r->numUnphased);
  return 0;
}

/* ------------------------------------------------------------------ ingest alone
 * The reference's readSeqFile + processHetPatterns(breakSymmetries = 1) per locus, exactly the sequence of
 * processAlignments (GPhoCS.c:258-440) without creating LocusData; what would be handed to initializeLocusData is
 * appended to the recorder.  sampleNames: one entry per slot, "" for the second slot of a diploid.  May be called
 * repeatedly.  Returns the number of loci or -1 (the reference's error text goes to stderr). */
static void clearRecorded(void) {
  for (int i = 0; i < numRecorded; i++) { free(recorded[i].chars); free(recorded[i].numPhases); free(recorded[i].counts); }
  numRecorded = 0;
}
int refh_ingest(const char *seqFile, int numSamples, char **sampleNames, int numLociToRead) {
  clearRecorded();
  if (readSeqFile(seqFile, numSamples, sampleNames, numLociToRead) < 0) return -1;
  int maxPhased = 4 * AlignmentData.numPatterns + 4;
  char **patt = (char **)malloc(sizeof(char *) * (AlignmentData.numPatterns + 1));
  char **phased = (char **)malloc(sizeof(char *) * maxPhased);
  phased[0] = (char *)malloc((size_t)numSamples * maxPhased);
  int *numPhases = (int *)malloc(sizeof(int) * maxPhased);
  int saveLeaves = recordLeaves;
  recordLeaves = numSamples;
  int rc = AlignmentData.numLoci;
  for (int gen = 0; gen < AlignmentData.numLoci && rc >= 0; gen++) {
    LocusProfile *lp = &AlignmentData.locusProfiles[gen];
    for (int p = 0; p < lp->numPatterns; p++) patt[p] = AlignmentData.patternArray[lp->patternIds[p]];
    int P = processHetPatterns(patt, lp->patternCounts, lp->numPatterns, 1, &phased, &numPhases, &maxPhased);
    if (P < 0) { printAlignmentError(); rc = -1; break; }
    if (numRecorded == capRecorded) {
      capRecorded = capRecorded ? 2 * capRecorded : 1024;
      recorded = (RecordedLocus *)realloc(recorded, capRecorded * sizeof(RecordedLocus));
    }
    RecordedLocus *r = &recorded[numRecorded++];
    r->locus = NULL;
    r->numLeaves = numSamples;
    r->numPatterns = P;
    r->numUnphased = lp->numPatterns;
    r->chars = (char *)malloc((size_t)(P > 0 ? P : 1) * numSamples);
    r->numPhases = (int *)malloc(sizeof(int) * (P > 0 ? P : 1));
    r->counts = (int *)malloc(sizeof(int) * (lp->numPatterns > 0 ? lp->numPatterns : 1));
    for (int p = 0; p < P; p++) { memcpy(r->chars + (size_t)p * numSamples, phased[p], numSamples); r->numPhases[p] = numPhases[p]; }
    for (int p = 0; p < lp->numPatterns; p++) r->counts[p] = lp->patternCounts[p];
  }
  recordLeaves = saveLeaves;
  free(patt); free(phased[0]); free(phased); free(numPhases);
  freeAlignmentData();
  return rc;
}

/* ------------------------------------------------------------------ set-up, as main() does it */
int refh_setup(const char *ctl, int nthreads, int verboseFlag) {
  int res;
  starttime();
  debug = 0;
  verbose = verboseFlag;
  if (nthreads < 1) nthreads = 1;
  omp_set_num_threads(nthreads);
  initGeneralInfo();
  res = readControlFile((char *)ctl);
  if (res != 0) return -1;
  if (dataSetup.popTree->numCurPops > NSPECIES || dataSetup.popTree->numMigBands > MAX_MIG_BANDS)
    return -2;
  res = checkSettings();
  finalizeNumParameters();
  if (res > 0) return -3;
  if (mcmcSetup.randomSeed < 0) mcmcSetup.randomSeed = 4242;
  if (!mcmcSetup.useData) return -4;
  res = processAlignments();
  if (res < 0) return -5;
  if (dataSetup.numSamples > NS) return -6;
  allocateAllMemory();
  initRandomGenerator(dataSetup.numLoci, mcmcSetup.randomSeed);
  fflush(stdout);
  return 0;
}

/* whole MCMC as configured in the control file (initializeMCMC + mcmc-iterations steps) */
int refh_run_mcmc(void) {
  int r = performMCMC();
  if (ioSetup.traceFile) fflush(ioSetup.traceFile);
  fflush(stdout);
  return r;
}
/* initial state only (random genealogies + first full likelihoods), GPhoCS.c:1122-1225 */
int refh_init_only(void) {
  int r = initializeMCMC();
  fflush(stdout);
  return r;
}
void refh_set_iterations(int n) { mcmcSetup.numSamples = n; }
void refh_set_threads(int n) { omp_set_num_threads(n < 1 ? 1 : n); }
int refh_max_threads(void) { return omp_get_max_threads(); }

/* ------------------------------------------------------------------ dimensions */
int refh_num_loci(void) { return dataSetup.numLoci; }
int refh_num_leaves(void) { return dataSetup.numSamples; }
int refh_num_pops(void) { return dataSetup.popTree->numPops; }
int refh_num_cur_pops(void) { return dataSetup.popTree->numCurPops; }
int refh_num_bands(void) { return dataSetup.popTree->numMigBands; }
int refh_root_pop(void) { return dataSetup.popTree->rootPop; }
void *refh_locus(int gen) { return dataState.lociData[gen]; }

/* ------------------------------------------------------------------ population tree */
void refh_get_pops(double *theta, double *age, double *sampleAge, int *father, int *son0, int *son1,
                   int *numSamplesPerPop) {
  PopulationTree *pt = dataSetup.popTree;
  for (int p = 0; p < pt->numPops; p++) {
    theta[p] = pt->pops[p]->theta;
    age[p] = pt->pops[p]->age;
    sampleAge[p] = pt->pops[p]->sampleAge;
    father[p] = pt->pops[p]->father ? pt->pops[p]->father->id : -1;
    son0[p] = (p >= pt->numCurPops) ? pt->pops[p]->sons[0]->id : -1;
    son1[p] = (p >= pt->numCurPops) ? pt->pops[p]->sons[1]->id : -1;
    numSamplesPerPop[p] = (p < pt->numCurPops) ? dataSetup.numSamplesPerPop[p] : 0;
  }
}
void refh_get_bands(int *source, int *target, double *rate, double *start, double *end) {
  PopulationTree *pt = dataSetup.popTree;
  for (int b = 0; b < pt->numMigBands; b++) {
    source[b] = pt->migBands[b].sourcePop;
    target[b] = pt->migBands[b].targetPop;
    rate[b] = pt->migBands[b].migRate;
    start[b] = pt->migBands[b].startTime;
    end[b] = pt->migBands[b].endTime;
  }
}
void refh_set_theta(int pop, double v) { dataSetup.popTree->pops[pop]->theta = v; }
void refh_set_mig_rate(int band, double v) { dataSetup.popTree->migBands[band].migRate = v; }

/* ------------------------------------------------------------------ per-locus genealogy */
void refh_get_tree(int gen, int *father, int *left, int *right, double *age, int *root, double *rate) {
  LocusData *ld = dataState.lociData[gen];
  int N = 2 * dataSetup.numSamples - 1;
  for (int i = 0; i < N; i++) {
    father[i] = getNodeFather(ld, i);
    left[i] = getNodeSon(ld, i, 0);
    right[i] = getNodeSon(ld, i, 1);
    age[i] = getNodeAge(ld, i);
  }
  *root = getLocusRoot(ld);
  *rate = getLocusMutationRate(ld);
}
void refh_get_node_pops(int gen, int *pops) {
  int N = 2 * dataSetup.numSamples - 1;
  for (int i = 0; i < N; i++) pops[i] = nodePops[gen][i];
}
/* living migration nodes of a locus: returns count; arrays sized MAX_MIGS */
int refh_get_migs(int gen, int *branch, int *band, int *targetPop, int *sourcePop, double *age) {
  int k = genetree_migs[gen].num_migs;
  for (int i = 0; i < k; i++) {
    int m = genetree_migs[gen].living_mignodes[i];
    branch[i] = genetree_migs[gen].mignodes[m].gtree_branch;
    band[i] = genetree_migs[gen].mignodes[m].migration_band;
    targetPop[i] = genetree_migs[gen].mignodes[m].target_pop;
    sourcePop[i] = genetree_migs[gen].mignodes[m].source_pop;
    age[i] = genetree_migs[gen].mignodes[m].age;
  }
  return k;
}
double refh_data_lnl(int gen) { return getLocusDataLikelihood(dataState.lociData[gen]); }
double refh_compute_data_lnl(int gen, int useOld) {
  return computeLocusDataLikelihood(dataState.lociData[gen], (unsigned short)useOld);
}
double refh_total_data_lnl(void) { return dataState.dataLogLikelihood; }
double refh_stored_gen_lnl(int gen) { return locus_data[gen].genLogLikelihood; }

/* ------------------------------------------------------------------ event chains, flattened
 * For every population (in id order) the chain is walked from first_event (patch.h:166-172) and each
 * event is emitted as (type, id, elapsed_time, num_lineages).  `id` is the migration BAND for IN_MIG
 * (resolved through genetree_migs, as recalcStats does at patch.c:2425), the band for
 * MIG_BAND_START/END, the node for COAL, and node_id otherwise.  popStart[Q+1] delimits chains.
 * Returns the number of events (call with NULL arrays to size). */
int refh_flatten_events(int gen, int *popStart, int *type, int *id, double *elapsed, int *numLineages) {
  int Q = dataSetup.popTree->numPops, k = 0;
  for (int pop = 0; pop < Q; pop++) {
    if (popStart) popStart[pop] = k;
    for (int ev = event_chains[gen].first_event[pop]; ev >= 0; ev = event_chains[gen].events[ev].next) {
      if (type) {
        Event *e = &event_chains[gen].events[ev];
        type[k] = (int)e->type;
        id[k] = (e->type == IN_MIG || e->type == OUT_MIG)
                    ? genetree_migs[gen].mignodes[e->node_id].migration_band
                    : e->node_id;
        elapsed[k] = e->elapsed_time;
        numLineages[k] = e->num_lineages;
      }
      k++;
    }
  }
  if (popStart) popStart[Q] = k;
  return k;
}
void refh_get_stats(int gen, double *coal_stats, int *num_coals, double *mig_stats, int *num_migs) {
  int Q = dataSetup.popTree->numPops, B = dataSetup.popTree->numMigBands;
  for (int p = 0; p < Q; p++) {
    coal_stats[p] = genetree_stats[gen].coal_stats[p];
    num_coals[p] = genetree_stats[gen].num_coals[p];
  }
  for (int b = 0; b < B; b++) {
    mig_stats[b] = genetree_stats[gen].mig_stats[b];
    num_migs[b] = genetree_stats[gen].num_migs[b];
  }
}
void refh_get_total_stats(double *coal_stats, int *num_coals, double *mig_stats, int *num_migs) {
  int Q = dataSetup.popTree->numPops, B = dataSetup.popTree->numMigBands;
  for (int p = 0; p < Q; p++) {
    coal_stats[p] = genetree_stats_total.coal_stats[p];
    num_coals[p] = genetree_stats_total.num_coals[p];
  }
  for (int b = 0; b < B; b++) {
    mig_stats[b] = genetree_stats_total.mig_stats[b];
    num_migs[b] = genetree_stats_total.num_migs[b];
  }
}
/* full recompute through the reference: computeGenetreeStats (patch.c:2330) + gtreeLnLikelihood (:2702) */
double refh_recompute_gen(int gen) {
  computeGenetreeStats(gen);
  return gtreeLnLikelihood(gen);
}
double refh_gen_lnl(int gen) { return gtreeLnLikelihood(gen); }
int refh_compute_total_stats(void) { return computeTotalStats(); }
int refh_check_all(void) { return checkAll(); }

/* ------------------------------------------------------------------ stand-alone LocusData helpers */
/* builds a GenericBinaryTree view over caller arrays and hands it to copyGenericTreeToLocus (.c:1023) */
int refh_set_tree(void *locus, int numLeaves, int *father, int *left, int *right, double *age, int root) {
  GenericBinaryTree t;
  t.numLeaves = numLeaves;
  t.rootId = root;
  t.leafNames = NULL;
  t.father = father;
  t.leftSon = left;
  t.rightSon = right;
  t.label1 = age;
  t.label2 = NULL;
  return copyGenericTreeToLocus((LocusData *)locus, &t);
}
/* initializeLocusData from a flat [numPatterns][numLeaves] char array (goes through the recorder too) */
int refh_init_locus(void *locus, int numLeaves, const char *chars, int numPatterns, int *numPhases,
                    int *counts) {
  char **rows = (char **)malloc(sizeof(char *) * (numPatterns > 0 ? numPatterns : 1));
  for (int p = 0; p < numPatterns; p++) rows[p] = (char *)chars + (size_t)p * numLeaves;
  int save = recordLeaves;
  recordLeaves = numLeaves;
  int r = initializeLocusData((LocusData *)locus, rows, numPatterns, numPhases, counts);
  recordLeaves = save;
  free(rows);
  return r;
}
void refh_get_locus_tree(void *locus, int numLeaves, int *father, int *left, int *right, double *age,
                         int *root) {
  LocusData *ld = (LocusData *)locus;
  for (int i = 0; i < 2 * numLeaves - 1; i++) {
    father[i] = getNodeFather(ld, i);
    left[i] = getNodeSon(ld, i, 0);
    right[i] = getNodeSon(ld, i, 1);
    age[i] = getNodeAge(ld, i);
  }
  *root = getLocusRoot(ld);
}

/* ------------------------------------------------------------------ CPU baseline timing loops */
static double now_s(void) {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_sec + 1e-6 * tv.tv_usec;
}
/* `reps` passes of computeLocusDataLikelihood(locus,0)+resetSaved over all loci, OpenMP static schedule
 * like the reference's own loops (MultiCoreUtils.h:8).  Returns best seconds per pass; *sum gets the
 * summed lnL of the last pass so the work cannot be optimised away. */
double refh_time_data_full(int reps, double *sum) {
  int L = dataSetup.numLoci;
  double best = 1e300, s = 0.0;
  for (int r = 0; r < reps; r++) {
    double t0 = now_s();
    s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int gen = 0; gen < L; gen++) {
      s += computeLocusDataLikelihood(dataState.lociData[gen], 0);
      resetSaved(dataState.lociData[gen]);
    }
    double t = now_s() - t0;
    if (t < best) best = t;
  }
  if (sum) *sum = s;
  return best;
}
/* same for the genealogy side: computeGenetreeStats + gtreeLnLikelihood, all loci */
double refh_time_gen_full(int reps, double *sum) {
  int L = dataSetup.numLoci;
  double best = 1e300, s = 0.0;
  for (int r = 0; r < reps; r++) {
    double t0 = now_s();
    s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int gen = 0; gen < L; gen++) {
      computeGenetreeStats(gen);
      s += gtreeLnLikelihood(gen);
    }
    double t = now_s() - t0;
    if (t < best) best = t;
  }
  if (sum) *sum = s;
  return best;
}
/* one timed pass of both (the unit bench.py's `--impl reference` reports) */
double refh_time_both_once(double *sumData, double *sumGen) {
  int L = dataSetup.numLoci;
  double sd = 0.0, sg = 0.0;
  double t0 = now_s();
#pragma omp parallel for schedule(static) reduction(+ : sd, sg)
  for (int gen = 0; gen < L; gen++) {
    sd += computeLocusDataLikelihood(dataState.lociData[gen], 0);
    resetSaved(dataState.lociData[gen]);
    computeGenetreeStats(gen);
    sg += gtreeLnLikelihood(gen);
  }
  double t = now_s() - t0;
  if (sumData) *sumData = sd;
  if (sumGen) *sumGen = sg;
  return t;
}
