/*
 * gphocs_oracle.c — CPU restatement of the G-PhoCS per-locus likelihood hot path.
 * TEST INFRASTRUCTURE ONLY — see gphocs_oracle.h for who may use it and how it is pinned.
 *
 * Written from the semantics of the reference, not from its text: the
 * reference keeps two node structs per node and swaps pointers; here the genealogy is a pair of
 * plain arrays (current / saved), each internal node owns two conditional-likelihood buffers and a
 * one-bit selector, and "flip" toggles the selector.  Arithmetic follows the reference operation
 * by operation (same association order, no FMA contraction: built with -ffp-contract=off) so the
 * log-likelihoods agree bit-for-bit with the compiled reference on the same inputs.
 */
#include "gphocs_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct OrcLocus {
  int n, N;            /* leaves, nodes */
  int P;               /* phased patterns */
  int live;            /* numLivePatterns (LocusDataLikelihood.c:294-298) */
  int *numPhases;      /* [P] 2^k on the first column of a phase group, 0 elsewhere */
  int *count;          /* [P] site multiplicity, meaningful where numPhases > 0 */
  double rate, lnL, savedLnL;
  int root, savedRoot; /* savedRoot = -1: unchanged */
  int copyAll;
  /* genealogy, current and saved */
  int *father, *left, *right, *svFather, *svLeft, *svRight;
  double *age, *svAge;
  /* conditional likelihoods: buf[(2*node + which)*4*P + 4*p + base]; sel[node] = current buffer */
  double *buf;
  unsigned char *sel, *recalc;
  int *changedNodes, numChangedNodes;
  int *changedConds, numChangedConds;
};

static double *clv(const OrcLocus *lc, int node, int saved) {
  int which = lc->sel[node] ^ (saved ? 1 : 0);
  return lc->buf + ((size_t)(2 * node + which)) * 4 * (size_t)(lc->P > 0 ? lc->P : 1);
}

/* resetSaved, LocusDataLikelihood.c:852-864 */
int orc_reset(OrcLocus *lc) {
  lc->copyAll = 0;
  lc->numChangedNodes = 0;
  lc->numChangedConds = 0;
  lc->savedRoot = -1;
  lc->savedLnL = lc->lnL;
  memset(lc->recalc, 0, lc->N);
  return 0;
}

/* createLocusData, LocusDataLikelihood.c:142-227 */
OrcLocus *orc_create(int numLeaves) {
  OrcLocus *lc = (OrcLocus *)calloc(1, sizeof(OrcLocus));
  if (!lc) return NULL;
  int N = 2 * numLeaves - 1;
  lc->n = numLeaves;
  lc->N = N;
  lc->rate = 1.0;
  lc->root = -1;
  lc->father = (int *)malloc(sizeof(int) * 6 * N);
  lc->left = lc->father + N;
  lc->right = lc->left + N;
  lc->svFather = lc->right + N;
  lc->svLeft = lc->svFather + N;
  lc->svRight = lc->svLeft + N;
  lc->age = (double *)calloc(2 * N, sizeof(double));
  lc->svAge = lc->age + N;
  lc->sel = (unsigned char *)calloc(2 * N, 1);
  lc->recalc = lc->sel + N;
  lc->changedNodes = (int *)malloc(sizeof(int) * 2 * N * 2);
  lc->changedConds = lc->changedNodes + 2 * N;
  for (int i = 0; i < 6 * N; i++) lc->father[i] = -1;
  orc_reset(lc);
  return lc;
}

/* initializeLocusData + computeLeafConditionals, LocusDataLikelihood.c:239-303, 1321-1398 */
int orc_init(OrcLocus *lc, const char *chars, int numPatterns, const int *numPhases, const int *counts) {
  if (!lc) return -1;
  int n = lc->n, N = lc->N, P = numPatterns;
  lc->P = 0;
  lc->buf = (double *)calloc((size_t)2 * N * 4 * (P > 0 ? P : 1), sizeof(double));
  lc->numPhases = (int *)malloc(sizeof(int) * 2 * (P > 0 ? P : 1));
  lc->count = lc->numPhases + (P > 0 ? P : 1);
  if (!lc->buf || !lc->numPhases) return -1;
  int unphased = 0;
  for (int p = 0; p < P; p++) {
    lc->numPhases[p] = numPhases[p];
    lc->count[p] = (counts && numPhases[p] > 0) ? counts[unphased++] : 0;
    for (int leaf = 0; leaf < n; leaf++) {
      double v[4] = {0.0, 0.0, 0.0, 0.0};
      switch (chars[(size_t)p * n + leaf]) {
        case 'T': v[0] = 1.0; break;
        case 'C': v[1] = 1.0; break;
        case 'A': v[2] = 1.0; break;
        case 'G': v[3] = 1.0; break;
        case 'N': v[0] = v[1] = v[2] = v[3] = 1.0; break;
        default:
          fprintf(stderr, "oracle: unexpected character '%c' for leaf %d in pattern %d\n",
                  chars[(size_t)p * n + leaf], leaf, p);
          return -1;
      }
      /* both buffers of a leaf hold the same values for ever (.c:1388-1390) */
      for (int which = 0; which < 2; which++)
        memcpy(lc->buf + ((size_t)(2 * leaf + which) * P + p) * 4, v, sizeof(v));
    }
  }
  lc->P = P;
  lc->live = counts ? P : 0;
  return 0;
}

void orc_free(OrcLocus *lc) {
  if (!lc) return;
  free(lc->buf);
  free(lc->numPhases);
  free(lc->father);
  free(lc->age);
  free(lc->sel);
  free(lc->changedNodes);
  free(lc);
}

void orc_set_rate(OrcLocus *lc, double rate) { lc->rate = rate; }
double orc_get_rate(const OrcLocus *lc) { return lc->rate; }
double orc_get_lnl(const OrcLocus *lc) { return lc->lnL; }

/* copyGenericTreeToLocus, LocusDataLikelihood.c:1023-1035 */
int orc_set_tree(OrcLocus *lc, const int *father, const int *left, const int *right, const double *age, int root) {
  for (int i = 0; i < lc->N; i++) {
    lc->father[i] = father[i];
    lc->left[i] = left[i];
    lc->right[i] = right[i];
    lc->age[i] = age[i];
  }
  lc->root = root;
  return 0;
}
void orc_get_tree(const OrcLocus *lc, int *father, int *left, int *right, double *age, int *root) {
  for (int i = 0; i < lc->N; i++) {
    father[i] = lc->father[i];
    left[i] = lc->left[i];
    right[i] = lc->right[i];
    age[i] = lc->age[i];
  }
  *root = lc->root;
}
void orc_get_clv(const OrcLocus *lc, int node, int saved, double *out) {
  memcpy(out, clv(lc, node, saved), sizeof(double) * 4 * lc->P);
}

/* copyNodeConditionals, LocusDataLikelihood.c:1889-1906: flip the buffer selector once per proposal */
static int flip(OrcLocus *lc, int node) {
  if (lc->P <= 0 || lc->recalc[node]) return 1;
  lc->changedConds[lc->numChangedConds++] = node;
  lc->recalc[node] = 1;
  lc->sel[node] ^= 1;
  return 0;
}

/* copyNodeToSaved, LocusDataLikelihood.c:1864-1876 */
static void saveNode(OrcLocus *lc, int node, int recalcFlag) {
  if (recalcFlag) flip(lc, node);
  lc->changedNodes[lc->numChangedNodes++] = node;
  lc->svAge[node] = lc->age[node];
  lc->svFather[node] = lc->father[node];
  lc->svLeft[node] = lc->left[node];
  lc->svRight[node] = lc->right[node];
}

/* computeEdgeConditionalJC, LocusDataLikelihood.c:1831-1848 */
double orc_edge_prob(double edgeLength) {
  if (edgeLength < 1e-100) return 0.0;
  return ((1 - exp(-4 * edgeLength / 3.0)) / 4.0);
}

/* computeSubtreeConditionals_new, LocusDataLikelihood.c:1650-1673 */
static void foldChild(const double *son, double *parent, const double *e) {
  double s = 0.0;
  for (int b = 0; b < 4; b++) s += son[b];
  if (s >= 4) return; /* all-missing subtree contributes a factor of one */
  double q = s * e[0];
  for (int b = 0; b < 4; b++) parent[b] *= (q + son[b] * e[1]);
}

/* computeConditionalJC_new, LocusDataLikelihood.c:1559-1636 */
static int recompute(OrcLocus *lc, int node, int numLive, const int *live, int override) {
  if (node < lc->n) return lc->recalc[node] ? 100 : 0;
  int l = lc->left[node], r = lc->right[node];
  int res = recompute(lc, l, numLive, live, override);
  res = recompute(lc, r, numLive, live, override) + res;
  if (!override && !res && !lc->recalc[node]) return 0;
  if (!override) flip(lc, node);
  double eL[2], eR[2];
  eL[0] = orc_edge_prob(lc->rate * (lc->age[node] - lc->age[l]));
  eL[1] = 1 - 4.0 * eL[0];
  eR[0] = orc_edge_prob(lc->rate * (lc->age[node] - lc->age[r]));
  eR[1] = 1 - 4.0 * eR[0];
  double *dst = clv(lc, node, 0);
  const double *cl = clv(lc, l, 0), *cr = clv(lc, r, 0);
  for (int k = 0; k < numLive; k++) {
    int p = live[k];
    for (int b = 0; b < 4; b++) dst[4 * p + b] = 1.0;
    foldChild(cl + 4 * p, dst + 4 * p, eL);
    foldChild(cr + 4 * p, dst + 4 * p, eR);
  }
  return 1;
}

/* computeLocusDataLikelihood, LocusDataLikelihood.c:426-483 */
double orc_compute(OrcLocus *lc, int useOld) {
  if (lc->live == 0) return 0.0;
  if (!useOld)
    for (int node = lc->n; node < lc->N; node++) flip(lc, node);
  lc->savedLnL = lc->lnL;
  int *live = (int *)malloc(sizeof(int) * lc->P);
  int numLive = 0;
  for (int p = 0; p < lc->P; p++)
    if (lc->count[p] > 0)
      for (int ph = 0; ph < lc->numPhases[p]; ph++) live[numLive++] = p + ph;
  if (numLive != lc->live) {
    fprintf(stderr, "oracle: there should be %d live patterns and there are %d\n", lc->live, numLive);
    free(live);
    return NAN;
  }
  int res = recompute(lc, lc->root, numLive, live, !useOld);
  if (!res) {
    free(live);
    return lc->lnL;
  }
  lc->lnL = 0.0;
  const double *rootClv = clv(lc, lc->root, 0);
  for (int k = 0; k < numLive;) {
    int p = live[k];
    double prob = 0.0;
    int numConds = 4 * lc->numPhases[p];
    for (int j = 0; j < numConds; j++) prob += rootClv[4 * p + j];
    lc->lnL += log(prob / numConds) * lc->count[p];
    k += lc->numPhases[p];
  }
  free(live);
  return lc->lnL;
}

/* adjustGenNodeAge, LocusDataLikelihood.c:875-881 */
int orc_adjust_age(OrcLocus *lc, int node, double age) {
  saveNode(lc, node, 1);
  lc->age[node] = age;
  return 0;
}

/* scaleAllNodeAges, LocusDataLikelihood.c:895-917 */
double orc_scale_all(OrcLocus *lc, double factor) {
  double old = lc->lnL;
  lc->copyAll = 1;
  for (int i = 0; i < lc->n; i++) lc->svFather[i] = lc->father[i];
  for (int i = 0; i < lc->N; i++) orc_adjust_age(lc, i, factor * lc->age[i]);
  orc_compute(lc, 1);
  return lc->lnL - old;
}

/* executeGenSPR, LocusDataLikelihood.c:931-1012 */
int orc_spr(OrcLocus *lc, int sub, int target, double age) {
  int targetFather = lc->father[target];
  int father = lc->father[sub];
  int grandpa = lc->father[father];
  int sibling = lc->left[father] + lc->right[father] - sub;
  orc_adjust_age(lc, father, age);
  if (target == sibling || target == father) return 0;
  /* prune */
  saveNode(lc, sibling, 0);
  lc->father[sibling] = grandpa;
  if (grandpa >= 0) {
    saveNode(lc, grandpa, 1);
    if (lc->left[grandpa] == father) lc->left[grandpa] = sibling;
    else lc->right[grandpa] = sibling;
  }
  /* regraft */
  lc->father[father] = targetFather;
  lc->left[father] = sub;
  lc->right[father] = target;
  if (target != grandpa) saveNode(lc, target, 0);
  lc->father[target] = father;
  if (targetFather < 0) {
    lc->savedRoot = target;
    lc->root = father;
    return 1;
  }
  if (targetFather == sibling) flip(lc, targetFather);
  else if (targetFather != grandpa) saveNode(lc, targetFather, 1);
  if (lc->left[targetFather] == target) lc->left[targetFather] = father;
  else lc->right[targetFather] = father;
  if (grandpa < 0) {
    lc->savedRoot = father;
    lc->root = sibling;
    return 2;
  }
  return 0;
}

static void restoreNode(OrcLocus *lc, int i) {
  lc->age[i] = lc->svAge[i];
  lc->father[i] = lc->svFather[i];
  lc->left[i] = lc->svLeft[i];
  lc->right[i] = lc->svRight[i];
}

/* revertToSaved, LocusDataLikelihood.c:768-841 */
int orc_revert(OrcLocus *lc) {
  lc->lnL = lc->savedLnL;
  if (lc->savedRoot >= 0) {
    lc->root = lc->savedRoot;
    lc->savedRoot = -1;
  }
  if (lc->copyAll) {
    /* wholesale swap of the node arrays: every node returns to its saved struct; the buffer that was
     * current before the flip becomes current again */
    for (int i = 0; i < lc->N; i++) {
      restoreNode(lc, i);
      if (lc->recalc[i]) lc->sel[i] ^= 1;
      else if (lc->P <= 0) lc->sel[i] ^= 1; /* pointers travel with the structs when nothing was flipped */
    }
    orc_reset(lc);
    return 0;
  }
  if (lc->numChangedConds == 0 && lc->numChangedNodes == 0) return 0;
  for (int k = 0; k < lc->numChangedNodes; k++) {
    int i = lc->changedNodes[k];
    restoreNode(lc, i);
    if (lc->recalc[i]) {
      lc->sel[i] ^= 1;
      lc->recalc[i] = 0;
    }
  }
  for (int k = 0; k < lc->numChangedConds; k++) {
    int i = lc->changedConds[k];
    if (lc->recalc[i]) {
      lc->sel[i] ^= 1;
      lc->recalc[i] = 0;
    }
  }
  lc->numChangedNodes = 0;
  lc->numChangedConds = 0;
  return 0;
}

/* checkLocusDataLikelihood, LocusDataLikelihood.c:717-758 */
int orc_check(OrcLocus *lc) {
  orc_compute(lc, 0);
  int ok = (lc->lnL == lc->savedLnL || fabs(1 - lc->lnL / lc->savedLnL) < 0.000000001);
  orc_reset(lc);
  return ok;
}

/* ================================================================================================
 * Genealogy likelihood
 * ================================================================================================ */

typedef struct {
  int type, id, next, prev;
  double elapsed;
} OEvent;

/* createEvent + createEventBefore, patch.c:1707-1802: walk while elapsed < remaining, so a new event of
 * equal age lands before the existing one; the following event's interval shrinks by the same amount */
static int insertEvent(OEvent *ev, int *first, int *freeHead, const OrcPopTree *pt, int pop, double age) {
  double delta = age - pt->age[pop];
  if (delta < 0) return -1;
  if (pop != pt->rootPop && age > pt->age[pt->father[pop]] + 0.000001) return -1;
  int e = first[pop];
  for (; ev[e].type != ORC_END_CHAIN && ev[e].elapsed < delta; e = ev[e].next) delta -= ev[e].elapsed;
  if (ev[e].elapsed < delta) {
    if (ev[e].elapsed < delta - 0.000001) return -1;
    delta = ev[e].elapsed;
  }
  int prev = ev[e].prev, nw = (*freeHead)++;
  ev[nw].next = e;
  ev[nw].prev = prev;
  ev[nw].elapsed = delta;
  ev[nw].type = ORC_DUMMY;
  ev[e].prev = nw;
  ev[e].elapsed -= delta;
  if (prev < 0) first[pop] = nw;
  else ev[prev].next = nw;
  return nw;
}

/* constructEventChain, patch.c:1961-2125 */
int orc_construct_events(const OrcPopTree *pt, int numLeaves, const int *nodePop, const double *nodeAge,
                         int numMigs, const int *migBand, const int *migTarget, const int *migSource,
                         const double *migAge, int *popStart, int *type, int *id, double *elapsed) {
  int Q = pt->numPops;
  int cap = Q + 2 * pt->numBands + pt->numCurPops + 2 * numMigs + numLeaves + 4;
  OEvent *ev = (OEvent *)malloc(sizeof(OEvent) * cap);
  int *first = (int *)malloc(sizeof(int) * Q);
  int freeHead = Q, e, fail = 0;
  for (int pop = 0; pop < Q; pop++) {
    ev[pop].type = ORC_END_CHAIN;
    ev[pop].next = ev[pop].prev = -1;
    ev[pop].id = pop;
    ev[pop].elapsed = (pop == pt->rootPop) ? 999 - pt->age[pop] : pt->age[pt->father[pop]] - pt->age[pop];
    first[pop] = pop;
  }
  for (int b = 0; b < pt->numBands && !fail; b++) {
    if ((e = insertEvent(ev, first, &freeHead, pt, pt->bandTarget[b], pt->bandStart[b])) < 0) { fail = 1; break; }
    ev[e].type = ORC_MIG_BAND_START; ev[e].id = b;
    if ((e = insertEvent(ev, first, &freeHead, pt, pt->bandTarget[b], pt->bandEnd[b])) < 0) { fail = 1; break; }
    ev[e].type = ORC_MIG_BAND_END; ev[e].id = b;
  }
  for (int pop = 0; pop < pt->numCurPops && !fail; pop++) {
    if ((e = insertEvent(ev, first, &freeHead, pt, pop, pt->sampleAge[pop])) < 0) { fail = 1; break; }
    ev[e].type = ORC_SAMPLES_START; ev[e].id = pop;
  }
  for (int m = 0; m < numMigs && !fail; m++) {
    if ((e = insertEvent(ev, first, &freeHead, pt, migTarget[m], migAge[m])) < 0) { fail = 1; break; }
    ev[e].type = ORC_IN_MIG; ev[e].id = migBand[m];
    if ((e = insertEvent(ev, first, &freeHead, pt, migSource[m], migAge[m])) < 0) { fail = 1; break; }
    ev[e].type = ORC_OUT_MIG; ev[e].id = migBand[m];
  }
  for (int node = numLeaves; node < 2 * numLeaves - 1 && !fail; node++) {
    if ((e = insertEvent(ev, first, &freeHead, pt, nodePop[node], nodeAge[node])) < 0) { fail = 1; break; }
    ev[e].type = ORC_COAL; ev[e].id = node;
  }
  int k = 0;
  if (!fail)
    for (int pop = 0; pop < Q; pop++) {
      popStart[pop] = k;
      for (e = first[pop]; e >= 0; e = ev[e].next, k++) {
        type[k] = ev[e].type; id[k] = ev[e].id; elapsed[k] = ev[e].elapsed;
      }
    }
  popStart[Q] = k;
  free(ev);
  free(first);
  return fail ? -1 : k;
}

/* populationPostOrder, patch.c:1936-1951 */
static int postOrder(const OrcPopTree *pt, int pop, int *out) {
  if (pop < pt->numCurPops) { out[0] = pop; return 1; }
  int size = postOrder(pt, pt->son0[pop], out);
  size += postOrder(pt, pt->son1[pop], out + size);
  out[size] = pop;
  return size + 1;
}

/* computeGenetreeStats (patch.c:2330-2354) + recalcStats (patch.c:2387-2513) */
int orc_gen_stats(const OrcPopTree *pt, const int *popStart, const int *type, const int *id,
                  const double *elapsed, int *numLineages, double *coal_stats, int *num_coals,
                  double *mig_stats, int *num_migs) {
  int Q = pt->numPops, B = pt->numBands;
  int *queue = (int *)malloc(sizeof(int) * (Q + B + 1)), *live = queue + Q, numLive;
  int *endLineages = (int *)malloc(sizeof(int) * Q);
  int rc = 0;
  postOrder(pt, pt->rootPop, queue);
  for (int b = 0; b < B; b++) { mig_stats[b] = 0.0; num_migs[b] = 0; }
  for (int i = 0; i < Q; i++) {
    int pop = queue[i];
    int n = (pop >= pt->numCurPops) ? endLineages[pt->son0[pop]] + endLineages[pt->son1[pop]] : 0;
    double coal = 0.0;
    int ncoal = 0;
    numLive = 0;
    for (int k = popStart[pop]; k < popStart[pop + 1]; k++) {
      double t = elapsed[k];
      numLineages[k] = n;
      coal += n * (n - 1) * t;
      for (int j = 0; j < numLive; j++) mig_stats[live[j]] += n * t;
      switch (type[k]) {
        case ORC_SAMPLES_START: n += pt->samplesPerPop[pop]; break;
        case ORC_COAL: ncoal++; n--; break;
        case ORC_IN_MIG: num_migs[id[k]]++; n--; break;
        case ORC_OUT_MIG: n++; break;
        case ORC_MIG_BAND_START: live[numLive++] = id[k]; num_migs[id[k]] = 0; mig_stats[id[k]] = 0.0; break;
        case ORC_MIG_BAND_END: {
          int j = 0;
          for (; j < numLive; j++) if (live[j] == id[k]) break;
          if (j == numLive) rc = -1; else live[j] = live[--numLive];
          break;
        }
        case ORC_DUMMY: case ORC_END_CHAIN: break;
        default: rc = -1;
      }
      if (type[k] == ORC_END_CHAIN) endLineages[pop] = n;
    }
    if (numLive != 0) rc = -1;
    coal_stats[pop] = coal;
    num_coals[pop] = ncoal;
  }
  free(queue);
  free(endLineages);
  return rc;
}

/* gtreeLnLikelihood, patch.c:2702-2723 (heredity factor 1, no admixed samples) */
double orc_gen_lnl(const OrcPopTree *pt, const double *coal_stats, const int *num_coals,
                   const double *mig_stats, const int *num_migs) {
  double lnLd = 0;
  for (int pop = 0; pop < pt->numPops; pop++) {
    double theta = pt->theta[pop];
    lnLd += num_coals[pop] * log(2 / theta) - coal_stats[pop] / (theta);
  }
  for (int b = 0; b < pt->numBands; b++) {
    double m = pt->bandRate[b];
    if (m > 0.0) lnLd += num_migs[b] * log(m) - mig_stats[b] * m;
  }
  return lnLd;
}
