/* gphocs_device_mcmc.c — the reference host handing its MCMC to the device-resident update steps.
 *
 * This is the binding a G-PhoCS maintainer adds to run a control file on the fast path: a C file
 * compiled NEXT TO the unmodified reference sources (it includes their headers) and linked with libgphocs_b200.so in
 * place of LocusDataLikelihood.o / AlignmentProcessor.o.  Everything up to and including initializeMCMC
 * (GPhoCS.c:1122-1225) is the reference's own code — control file, alignments (readSeqFile comes from the library),
 * population tree, starting genealogies, event chains, first likelihoods through the LocusData call surface.  Then,
 * instead of performMCMC's loop (GPhoCS.c:1476-1690), the state is flattened once into gphocsSamplerCreate /
 * gphocsSamplerSetMigration / gphocsSamplerSetAncient and every iteration runs on the GPU; the library writes the trace
 * file in performMCMC's format.  GPhoCS.c is compiled with -Dmain=gphocs_reference_main so that this file's main()
 * is the program's (oracle/Makefile: devhost).
 *
 * Not taken over (the program says so and stops): admixture, find-finetunes, genetree-samples > 1.
 */
#include <getopt.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "utils.h"
#include "MCMCcontrol.h"
#include "AlignmentProcessor.h"
#include "GenericTree.h"
#include "PopulationTree.h"
#include "LocusDataLikelihood.h"
#include "MultiCoreUtils.h"
#include "GPhoCS.h"
#include "patch.h"

/* the part of include/gphocs_b200.h this file uses (the reference's own LocusDataLikelihood.h is already included,
 * so the product header, which repeats that call surface, is not) */
typedef struct GphocsStore GphocsStore;
typedef struct GphocsSampler GphocsSampler;
GphocsStore *gpuLociStore(void);
GphocsSampler *gphocsSamplerCreate(GphocsStore *s, int numPops, int numCurPops, const int *popFather, const int *popSon0,
                                   const int *popSon1, const int *samplesPerPop, const double *theta, const double *tau,
                                   const double *thetaAlpha, const double *thetaBeta, const double *tauAlpha,
                                   const double *tauBeta, const int *nodePop, unsigned long long seed);
int gphocsSamplerDestroy(GphocsSampler *sm);
int gphocsSamplerSetFinetunes(GphocsSampler *sm, double coalTime, double theta, double tau, double mixing);
int gphocsSamplerSetMigFinetunes(GphocsSampler *sm, double migTime, double migRate);
int gphocsSamplerSetMigration(GphocsSampler *sm, int numBands, const int *bandSrc, const int *bandTgt, const double *migRate,
                              const double *migAlpha, const double *migBeta, const int *numMigs, const int *migBranch,
                              const int *migBand, const double *migAge);
int gphocsSamplerSetMigRates(GphocsSampler *sm, const double *migRate);
int gphocsSamplerSetAncient(GphocsSampler *sm, const int *estimate, const double *finetune, double locusRateFinetune,
                            double rateAlpha);
int gphocsSamplerIterate(GphocsSampler *sm, int iterations, double *trace);
int gphocsSamplerOpenTrace(GphocsSampler *sm, const char *path, const char *const *popNames, double thetaTauPrint,
                           double migRatePrint, int sampleSkip);
int gphocsSamplerCloseTrace(GphocsSampler *sm);
int gphocsSamplerGetState(GphocsSampler *sm, double *theta, double *tau, long long *accepted, long long *proposed);
int gphocsSamplerCheck(GphocsSampler *sm, double *maxStatErr, double *maxLnLErr);
int gphocsSamplerDownload(GphocsSampler *sm, int *nodePop);
long long gphocsKernelLaunchCount(void);

static void die(const char *what) {
  fprintf(stderr, "\nError: %s\n", what);
  exit(-1);
}

/* performMCMC's role: initializeMCMC, hand-over, iterations on the device, trace */
static int performDeviceMCMC(void) {
  PopulationTree *pt = dataSetup.popTree;
  const int Q = pt->numPops, C = pt->numCurPops, B = pt->numMigBands, L = dataSetup.numLoci;
  const int N = 2 * dataSetup.numSamples - 1;
  if (admixed_samples.number > 0) die("admixed samples are not taken over by the device-resident update steps");
  if (mcmcSetup.findFinetunes) die("find-finetunes is not taken over by the device-resident update steps: give the finetunes");
  if (mcmcSetup.genetreeSamples != 1) die("genetree-samples must be 1 for the device-resident update steps");

  printf("Starting MCMC on the GPU: %d burnin, %d running, sampled every %d iteration(s).\n", mcmcSetup.burnin,
         mcmcSetup.numSamples, mcmcSetup.sampleSkip + 1);
  const int totalCoals = initializeMCMC();   /* starting genealogies, event chains, first likelihoods (GPhoCS.c:1122) */
  if (totalCoals <= 0) {
    printf("Error while initializing MCMC.\n");
    return -1;
  }

  /* ---- the model (PopulationTree.h:80-125) */
  int father[2 * NSPECIES], son0[2 * NSPECIES], son1[2 * NSPECIES], samples[2 * NSPECIES], estimate[2 * NSPECIES];
  double theta[2 * NSPECIES], tau[2 * NSPECIES], thA[2 * NSPECIES], thB[2 * NSPECIES], tauA[2 * NSPECIES], tauB[2 * NSPECIES];
  double ftSampleAge[2 * NSPECIES];
  const char *names[2 * NSPECIES];
  double ftTau = -1.0;
  int anySampleAge = 0;
  for (int p = 0; p < Q; p++) {
    const Population *pop = pt->pops[p];
    father[p] = pop->father ? pop->father->id : -1;
    son0[p] = p < C ? -1 : pop->sons[0]->id;
    son1[p] = p < C ? -1 : pop->sons[1]->id;
    samples[p] = p < C ? dataSetup.numSamplesPerPop[p] : 0;
    theta[p] = pop->theta;
    tau[p] = p < C ? pop->sampleAge : pop->age;      /* a current population's "tau" is the age of its samples */
    thA[p] = pop->thetaPrior.alpha; thB[p] = pop->thetaPrior.beta;
    tauA[p] = pop->agePrior.alpha; tauB[p] = pop->agePrior.beta;
    estimate[p] = p < C ? pop->updateSampleAge : 0;
    ftSampleAge[p] = mcmcSetup.finetunes.taus[p];
    anySampleAge |= estimate[p];
    names[p] = pop->name;
    if (p >= C) {
      if (ftTau < 0.0) ftTau = mcmcSetup.finetunes.taus[p];
      else if (mcmcSetup.finetunes.taus[p] != ftTau)
        fprintf(stderr, "Warning: the device-resident steps use one finetune for all split times (%g); %s asks for %g.\n", ftTau,
                pop->name, mcmcSetup.finetunes.taus[p]);
    }
  }
  /* ---- population of every genealogy node (nodePops, patch.h:123) */
  int *nodePop = (int *)malloc(sizeof(int) * (size_t)L * N);
  if (!nodePop) die("out of memory");
  for (int gen = 0; gen < L; gen++)
    for (int i = 0; i < N; i++) nodePop[(size_t)gen * N + i] = nodePops[gen][i];

  GphocsStore *S = gpuLociStore();           /* the loci the host created through createLocusData, now resident */
  GphocsSampler *M = gphocsSamplerCreate(S, Q, C, father, son0, son1, samples, theta, tau, thA, thB, tauA, tauB, nodePop,
                                         (unsigned long long)mcmcSetup.randomSeed);
  if (!M) die("the device-resident sampler could not be created");
  gphocsSamplerSetFinetunes(M, mcmcSetup.finetunes.coalTime, mcmcSetup.finetunes.theta, ftTau > 0.0 ? ftTau : 0.0,
                            mcmcSetup.doMixing ? mcmcSetup.finetunes.mixing : 0.0);
  /* ---- migration bands and the migration events of every genealogy (genetree_migs, patch.h:138-148) */
  double *migRate = NULL;
  if (B > 0) {
    int *src = (int *)malloc(sizeof(int) * B), *tgt = (int *)malloc(sizeof(int) * B);
    double *mA = (double *)malloc(sizeof(double) * B), *mB = (double *)malloc(sizeof(double) * B);
    migRate = (double *)malloc(sizeof(double) * B);
    int *numMigs = (int *)calloc((size_t)L, sizeof(int));
    int *br = (int *)calloc((size_t)L * MAX_MIGS, sizeof(int)), *bd = (int *)calloc((size_t)L * MAX_MIGS, sizeof(int));
    double *ag = (double *)calloc((size_t)L * MAX_MIGS, sizeof(double));
    if (!src || !tgt || !mA || !mB || !migRate || !numMigs || !br || !bd || !ag) die("out of memory");
    for (int b = 0; b < B; b++) {
      src[b] = pt->migBands[b].sourcePop; tgt[b] = pt->migBands[b].targetPop;
      migRate[b] = pt->migBands[b].migRate;        /* 0 until iteration start-mig (PopulationTree.c:393-397) */
      mA[b] = pt->migBands[b].migRatePrior.alpha; mB[b] = pt->migBands[b].migRatePrior.beta;
    }
    for (int gen = 0; gen < L; gen++) {
      numMigs[gen] = genetree_migs[gen].num_migs;
      for (int k = 0; k < genetree_migs[gen].num_migs; k++) {
        const int id = genetree_migs[gen].living_mignodes[k];
        br[(size_t)gen * MAX_MIGS + k] = genetree_migs[gen].mignodes[id].gtree_branch;
        bd[(size_t)gen * MAX_MIGS + k] = genetree_migs[gen].mignodes[id].migration_band;
        ag[(size_t)gen * MAX_MIGS + k] = genetree_migs[gen].mignodes[id].age;
      }
    }
    if (gphocsSamplerSetMigration(M, B, src, tgt, migRate, mA, mB, numMigs, br, bd, ag)) die("the migration bands were refused");
    gphocsSamplerSetMigFinetunes(M, mcmcSetup.finetunes.migTime, mcmcSetup.finetunes.migRate);
    free(src); free(tgt); free(mA); free(mB); free(numMigs); free(br); free(bd); free(ag);
  }
  /* ---- estimated sample ages and locus-rate variation (UpdateSampleAge GPhoCS.c:4006, UpdateLocusRate :4598) */
  if (anySampleAge || mcmcSetup.mutRateMode == 1)
    gphocsSamplerSetAncient(M, estimate, ftSampleAge, mcmcSetup.mutRateMode == 1 ? mcmcSetup.finetunes.locusRate : 0.0,
                            mcmcSetup.varRatesAlpha);

  /* ---- iterations: burn-in, then the sampled ones; migration rates are drawn after iteration start-mig (GPhoCS.c:1738) */
  const int logEvery = ioSetup.samplesPerLog > 0 ? ioSetup.samplesPerLog : 100;
  const double thetaPrint = mcmcSetup.printFactors[0];
  const double migPrint = B > 0 ? mcmcSetup.printFactors[2 * Q - C] : 1.0;
  const long long launches0 = gphocsKernelLaunchCount();
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  int traceOpen = 0, ratesDrawn = B == 0;
  for (int iteration = -mcmcSetup.burnin; iteration < mcmcSetup.numSamples;) {
    if (iteration >= 0 && !traceOpen) {
      if (gphocsSamplerOpenTrace(M, ioSetup.traceFileName, names, thetaPrint, migPrint, mcmcSetup.sampleSkip)) return -1;
      traceOpen = 1;
    }
    /* run up to the next point where the host has something to do: trace start, start-mig, log line, end */
    int stop = mcmcSetup.numSamples;
    if (iteration < 0) stop = 0;
    if (!ratesDrawn && mcmcSetup.startMig + 1 > iteration && mcmcSetup.startMig + 1 < stop) stop = mcmcSetup.startMig + 1;
    const int nextLog = (iteration / logEvery + 1) * logEvery;
    if (iteration >= 0 && nextLog < stop) stop = nextLog;
    if (gphocsSamplerIterate(M, stop - iteration, NULL)) die("an MCMC iteration failed on the GPU");
    iteration = stop;
    if (!ratesDrawn && iteration == mcmcSetup.startMig + 1) {
      sampleMigRates(pt);                                   /* the reference's own draw from the priors */
      for (int b = 0; b < B; b++) migRate[b] = pt->migBands[b].migRate;
      if (gphocsSamplerSetMigRates(M, migRate)) die("the migration rates were refused");
      ratesDrawn = 1;
    }
    if (iteration > 0 && iteration % logEvery == 0) {
      double se = 0.0, le = 0.0;
      const int bad = gphocsSamplerCheck(M, &se, &le);      /* checkAll at every log line (GPhoCS.c:1814) */
      if (bad != 0) {
        fprintf(stderr, "\nError:  --  Aborting when logging after MCMC iteration %d, due to data structure inconsistency.\n\n", iteration);
        exit(-1);
      }
      long long acc[10], prop[10];
      gphocsSamplerGetState(M, theta, tau, acc, prop);
      printf("%8d  coal-time %5.1f%%  SPR %5.1f%%  theta %5.1f%%  tau %5.1f%%  mixing %5.1f%%  mig-rate %5.1f%%  | max stat drift %.1e\n",
             iteration, 100.0 * acc[0] / fmax(1.0, (double)prop[0]), 100.0 * acc[1] / fmax(1.0, (double)prop[1]),
             100.0 * acc[2] / fmax(1.0, (double)prop[2]), 100.0 * acc[3] / fmax(1.0, (double)prop[3]),
             100.0 * acc[4] / fmax(1.0, (double)prop[4]), 100.0 * acc[5] / fmax(1.0, (double)prop[5]), se);
      fflush(stdout);
    }
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  const double secs = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  const int total = mcmcSetup.burnin + mcmcSetup.numSamples;
  printf("MCMC done: %d iterations in %.3f s (%.2f iterations/s), %lld kernel launches.\n", total, secs, total / fmax(secs, 1e-9),
         gphocsKernelLaunchCount() - launches0);
  /* ---- the final state back into the host's structures: parameters, genealogies (through the LocusData mirror) */
  gphocsSamplerGetState(M, theta, tau, NULL, NULL);
  for (int p = 0; p < Q; p++) {
    pt->pops[p]->theta = theta[p];
    if (p < C) pt->pops[p]->sampleAge = tau[p]; else pt->pops[p]->age = tau[p];
  }
  gphocsSamplerDownload(M, nodePop);
  for (int gen = 0; gen < L; gen++)
    for (int i = 0; i < N; i++) nodePops[gen][i] = nodePop[(size_t)gen * N + i];
  if (traceOpen) gphocsSamplerCloseTrace(M);
  gphocsSamplerDestroy(M);
  free(nodePop);
  free(migRate);
  return 0;
}

/* the reference's main (GPhoCS.c:49-249) with performMCMC replaced; options: -v, -n <threads> (host threads) */
int main(int argc, char *argv[]) {
  int c, numThreads = -1;
  starttime();
  debug = 0;
  while ((c = getopt(argc, argv, "hvn:")) != -1) {
    if (c == 'v') verbose = 1;
    else if (c == 'n') numThreads = atoi(optarg);
    else if (c == 'h') { printf("Usage: %s <control-file-name> [secondary-control-file-name] [-v] [-n threads]\n", argv[0]); exit(-1); }
  }
  if (argc <= 1 || argv[optind] == NULL) {
    printf("Usage: %s <control-file-name> [secondary-control-file-name] [-v] [-n threads]\n", argv[0]);
    exit(-1);
  }
  printf("G-PhoCS host " GPHOCS_VERSION_NUM " with the MCMC update steps on the GPU (libgphocs_b200)\n");
  omp_set_num_threads(numThreads > 0 ? numThreads : omp_get_max_threads());
  printf("Reading control settings from file %s...\n", argv[optind]);
  initGeneralInfo();
  if (readControlFile(argv[optind]) != 0) exit(-1);
  if (argv[optind + 1] != NULL && readSecondaryControlFile(argv[optind + 1]) != 0) exit(-1);
  if (dataSetup.popTree->numCurPops > NSPECIES) die("too many populations (NSPECIES, patch.h)");
  if (dataSetup.popTree->numMigBands > MAX_MIG_BANDS) die("too many migration bands (MAX_MIG_BANDS, patch.h)");
  const int errors = checkSettings();
  finalizeNumParameters();
  if (errors > 0) {
    fprintf(stderr, "Found %d errors when processing control settings.\n", errors);
    exit(-1);
  }
  if (mcmcSetup.randomSeed < 0) mcmcSetup.randomSeed = abs(2 * (int)time(NULL) + 1);
  if (verbose) printPriorSettings();
  if ((mcmcSetup.useData ? processAlignments() : initLociWithoutData()) < 0) exit(-1);
  if (dataSetup.numSamples > NS) die("too many samples (NS, patch.h)");
  allocateAllMemory();
  printf("\n");
  initRandomGenerator(dataSetup.numLoci, mcmcSetup.randomSeed);
  if (performDeviceMCMC() != 0) exit(-1);
  freeAlignmentData();
  freeAllMemory();
  return 0;
}
