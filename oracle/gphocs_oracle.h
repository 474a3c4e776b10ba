/*
 * gphocs_oracle.h — CPU restatement of the G-PhoCS per-locus likelihood hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library, and only as the checker.  The product library
 * (libgphocs_b200.so) never links, loads or calls anything in oracle/.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against outputs of the reference itself, compiled here from its own sources
 * into oracle/_ref/libgphocs_ref.so (oracle/Makefile): tests/test_oracle_vs_reference.py drives both
 * in lock-step, and tests/golden/ (npz files) holds vectors dumped from the reference by
 * tests/golden/make_golden.py for boxes where /root/reference does not exist.
 */
#ifndef GPHOCS_ORACLE_H
#define GPHOCS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcLocus OrcLocus;

/* ---- data likelihood P(X|G): follows src/LocusDataLikelihood.c (line refs at each definition) ---- */
OrcLocus *orc_create(int numLeaves);
int orc_init(OrcLocus *lc, const char *chars /* [numPatterns][numLeaves] */, int numPatterns,
             const int *numPhases, const int *counts /* one per phase group, may be NULL */);
void orc_free(OrcLocus *lc);
void orc_set_rate(OrcLocus *lc, double rate);
double orc_get_rate(const OrcLocus *lc);
int orc_set_tree(OrcLocus *lc, const int *father, const int *left, const int *right, const double *age, int root);
void orc_get_tree(const OrcLocus *lc, int *father, int *left, int *right, double *age, int *root);
double orc_compute(OrcLocus *lc, int useOldConditionals);
double orc_get_lnl(const OrcLocus *lc);
int orc_adjust_age(OrcLocus *lc, int node, double age);
double orc_scale_all(OrcLocus *lc, double factor);
int orc_spr(OrcLocus *lc, int subtreeRoot, int targetBranch, double age);
int orc_revert(OrcLocus *lc);
int orc_reset(OrcLocus *lc);
int orc_check(OrcLocus *lc);
/* current (saved=0) or saved (saved=1) conditional-likelihood vector of a node: out[numPatterns*4] */
void orc_get_clv(const OrcLocus *lc, int node, int saved, double *out);
double orc_edge_prob(double edgeLength);

/* ---- genealogy likelihood P(G|M): follows src/patch.c ---- */
enum { ORC_COAL = 0, ORC_IN_MIG, ORC_OUT_MIG, ORC_MIG_BAND_START, ORC_MIG_BAND_END, ORC_SAMPLES_START,
       ORC_END_CHAIN, ORC_DUMMY };

typedef struct {
  int numPops, numCurPops, numBands, rootPop;
  const double *theta;      /* [numPops] */
  const double *age;        /* [numPops] start time of each population (0 for current pops) */
  const double *sampleAge;  /* [numCurPops] */
  const int *father;        /* [numPops], -1 for the root population */
  const int *son0, *son1;   /* [numPops], -1 for current populations */
  const int *samplesPerPop; /* [numCurPops] haploid leaves */
  const int *bandSource, *bandTarget; /* [numBands] */
  const double *bandRate, *bandStart, *bandEnd;
} OrcPopTree;

/* constructEventChain (patch.c:1961-2125) for one genealogy, emitted flattened per population:
 * popStart[numPops+1], then (type,id,elapsed) per event in chain order. id = band for mig/band events,
 * node for COAL, pop for END_CHAIN/SAMPLES_START. Returns number of events or -1. */
int orc_construct_events(const OrcPopTree *pt, int numLeaves, const int *nodePop, const double *nodeAge,
                         int numMigs, const int *migBand, const int *migTarget, const int *migSource,
                         const double *migAge, int *popStart, int *type, int *id, double *elapsed);
/* computeGenetreeStats + recalcStats (patch.c:2330-2354, 2387-2513) over a flattened chain set.
 * Writes num_lineages per event and the four statistics arrays. Returns 0, or -1 on a malformed chain. */
int orc_gen_stats(const OrcPopTree *pt, const int *popStart, const int *type, const int *id,
                  const double *elapsed, int *numLineages, double *coal_stats, int *num_coals,
                  double *mig_stats, int *num_migs);
/* gtreeLnLikelihood (patch.c:2702-2723) */
double orc_gen_lnl(const OrcPopTree *pt, const double *coal_stats, const int *num_coals,
                   const double *mig_stats, const int *num_migs);

/* ---- pattern / phase producer: follows src/AlignmentProcessor.c, see ingest_oracle.c ---- */
int orc_base_type(char c);
int orc_canonize_column(const char *column, char *pattern, int n);
int orc_locus_patterns(const char *const *rows, int n, int seqLength, char *patterns, int *counts);
int orc_symmetry_breaks(const char *patterns, const int *counts, int U, int n, unsigned char *breaks);
int orc_expand_phases(const char *patterns, const int *counts, int U, int n, const unsigned char *isDiploid,
                      int breakSymmetries, char *phased, int *numPhases, int capacity);

#ifdef __cplusplus
}
#endif
#endif
