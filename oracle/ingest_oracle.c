/*
 * ingest_oracle.c — CPU restatement of the reference's pattern / phase producer.
 *
 * TEST INFRASTRUCTURE ONLY (see gphocs_oracle.h).  Pinned against the reference itself: oracle/ref_harness.c
 * interposes initializeLocusData and records exactly what processAlignments (GPhoCS.c:258-440) hands over for
 * every locus; tests/test_oracle_vs_reference.py compares those records with this file's output on the same
 * sequence files, tests/golden/ingest_*.npz carries them to boxes without /root/reference.
 *
 * Slots: the control file's sample list has one slot per haploid sample and two per diploid sample (the second
 * one nameless, AlignmentProcessor.c:219-226).  A column holds the sample's character in the first slot of a
 * diploid and 'N' in the second (:898-908).
 */
#include <stdlib.h>
#include <string.h>

#include "gphocs_oracle.h"

/* symbol order of the canonical alphabet (AlignmentProcessor.c:61): 4 bases, 6 two-way, 4 three-way codes, N */
static const char kSymbols[] = "TCAGYWKMSRVDBHN";

static int symbolIndex(char c) {
  const char *p = c ? strchr(kSymbols, c) : NULL;
  return p ? (int)(p - kSymbols) : -1;
}

/* 0 base, 1 partial ambiguity (two- or three-way), 2 N, -1 not a symbol  (getBaseType, :1467-1477) */
int orc_base_type(char c) {
  const int s = symbolIndex(c);
  if (s < 0) return -1;
  return s < 4 ? 0 : (s < 14 ? 1 : 2);
}

/* image of every symbol under each of the 24 base permutations (initializeBaseTransformations, :1518-1590):
 * a two-way code {a<b} has index 2a+b+3 (10 folded onto 9), a three-way code 10 + the excluded base */
static int permImage[24][15];
static int permReady = 0;

static int pairCode(int a, int b) {
  if (a > b) { const int t = a; a = b; b = t; }
  const int c = 2 * a + b + 3;
  return c == 10 ? 9 : c;
}

static void buildPermutations(void) {
  int k = 0;
  for (int a = 0; a < 4; a++)
    for (int b = 0; b < 4; b++)
      for (int c = 0; c < 4; c++) {
        if (a == b || a == c || b == c) continue;
        const int d = 6 - a - b - c;
        int *im = permImage[k++];
        im[0] = a; im[1] = b; im[2] = c; im[3] = d;
        for (int x = 0; x < 4; x++) {
          im[10 + x] = 10 + im[x];
          for (int y = x + 1; y < 4; y++) im[pairCode(x, y)] = pairCode(im[x], im[y]);
        }
        im[14] = 14;
      }
  permReady = 1;
}

/* cannonizeJCpattern (:1595-1655): symbol by symbol, the smallest image any still-consistent permutation offers */
int orc_canonize_column(const char *column, char *pattern, int n) {
  if (!permReady) buildPermutations();
  unsigned alive = 0xFFFFFFu;
  for (int s = 0; s < n; s++) {
    const int sym = symbolIndex(column[s]);
    if (sym < 0) return -1;
    int best = 100;
    for (int p = 0; p < 24; p++)
      if (((alive >> p) & 1u) && permImage[p][sym] < best) best = permImage[p][sym];
    for (int p = 0; p < 24; p++)
      if (permImage[p][sym] != best) alive &= ~(1u << p);
    pattern[s] = kSymbols[best];
  }
  return 0;
}

/* processLocusAlignment (:871-990): rows[slot] = the slot's sequence or NULL; columns of N only are dropped;
 * distinct canonical patterns in order of first appearance with their multiplicities.  Returns their number. */
int orc_locus_patterns(const char *const *rows, int n, int seqLength, char *patterns, int *counts) {
  char *column = (char *)malloc(2 * (size_t)n);
  char *canon = column + n;
  int U = 0;
  for (int site = 0; site < seqLength; site++) {
    int informative = 0;
    for (int s = 0; s < n; s++) {
      column[s] = rows[s] ? rows[s][site] : 'N';
      informative |= column[s] != 'N';
    }
    if (!informative) continue;
    if (orc_canonize_column(column, canon, n) < 0) { free(column); return -1; }
    int u = 0;
    while (u < U && memcmp(patterns + (size_t)u * n, canon, n) != 0) u++;
    if (u == U) {
      memcpy(patterns + (size_t)U * n, canon, n);
      counts[U++] = 0;
    }
    counts[u]++;
  }
  free(column);
  return U;
}

/* computeHetSymmetryBreaks (:1706-1895).  Greedy: only patterns seen once are eligible; the pattern with the
 * highest score (2^hets at the start, halved each time the pattern itself is chosen) gives up its LAST live het,
 * which is then arbitrarily phased there and retired from every other eligible pattern (swap-with-last removal,
 * patterns left without live hets leave the list the same way).  breaks[u*n + slot] = 1 where phase is fixed. */
int orc_symmetry_breaks(const char *patterns, const int *counts, int U, int n, unsigned char *breaks) {
  int *hets = (int *)malloc(sizeof(int) * ((size_t)U * n + 3 * (size_t)U + 1));
  int *numHets = hets + (size_t)U * n, *live = numHets + U, *where = live + U;
  double *score = (double *)malloc(sizeof(double) * (size_t)(U + 1));
  memset(breaks, 0, (size_t)U * n);
  int numLive = 0, chosen = -1;
  double top = -1.0;
  for (int u = 0; u < U; u++) {
    numHets[u] = 0;
    where[u] = -1;
    score[u] = -1.0;
    if (counts[u] > 1) continue;
    for (int s = 0; s < n; s++)
      if (orc_base_type(patterns[(size_t)u * n + s]) == 1) hets[(size_t)u * n + numHets[u]++] = s;
    if (numHets[u] > 0) {
      score[u] = (double)(1u << (numHets[u] < 30 ? numHets[u] : 30));
      for (int k = 30; k < numHets[u]; k++) score[u] *= 2.0;
      where[u] = numLive;
      live[numLive++] = u;
    }
    if (top < score[u]) { top = score[u]; chosen = u; }
  }
  while (top > 0.0) {
    const int slot = hets[(size_t)chosen * n + --numHets[chosen]];
    breaks[(size_t)chosen * n + slot] = 1;
    score[chosen] = numHets[chosen] <= 0 ? -1.0 : score[chosen] / 2.0;
    top = score[chosen];
    for (int i = 0; i < numLive;) {
      const int u = live[i];
      int *h = hets + (size_t)u * n;
      for (int k = 0; k < numHets[u]; k++)
        if (h[k] == slot) { h[k] = h[--numHets[u]]; break; }
      if (numHets[u] > 0) {
        i++;
      } else {
        numLive--;
        live[where[u]] = live[numLive];
        where[live[where[u]]] = where[u];
        where[u] = -1;
        score[u] = -1.0;
      }
      if (top < score[u]) { top = score[u]; chosen = u; }
    }
  }
  free(hets);
  free(score);
  return 0;
}

/* translateAmbiguity (:2298-2340): the two bases of a diploid genotype; three-way codes and N give N,N */
static void genotypeBases(char c, char *out) {
  static const char *two[] = {"YTC", "KTG", "WTA", "SCG", "MAC", "RAG"};
  out[0] = out[1] = 'N';
  if (c == 'T' || c == 'C' || c == 'A' || c == 'G') { out[0] = out[1] = c; return; }
  for (int k = 0; k < 6; k++)
    if (two[k][0] == c) { out[0] = two[k][1]; out[1] = two[k][2]; }
}

/* processHetPatterns + getAllPhases (:998-1158, :2242-2290).  Every partial-ambiguity genotype that is not
 * arbitrarily phased doubles the number of columns of its pattern; getAllPhases steps through the phasings like a
 * binary counter over the free genotypes in slot order, so column k swaps the two bases of free genotype j iff bit j
 * of k is set.  numPhases: the count on the first column of a pattern, 0 on the others.
 * Returns the number of phased columns P, or -(P) if `capacity` columns do not hold them (nothing written). */
int orc_expand_phases(const char *patterns, const int *counts, int U, int n, const unsigned char *isDiploid,
                      int breakSymmetries, char *phased, int *numPhases, int capacity) {
  unsigned char *breaks = (unsigned char *)malloc((size_t)U * n + 1);
  if (orc_symmetry_breaks(patterns, counts, U, n, breaks) < 0) { free(breaks); return -1; }
  long total = 0;
  for (int u = 0; u < U; u++) {
    long c = 1;
    for (int s = 0; s < n; s++)
      if (orc_base_type(patterns[(size_t)u * n + s]) == 1 && !(breakSymmetries && breaks[(size_t)u * n + s])) c *= 2;
    total += c;
  }
  if (total > capacity) { free(breaks); return (int)-total; }
  char *base = (char *)malloc((size_t)n + 2);
  int *freeSlot = (int *)malloc(sizeof(int) * (size_t)(n + 1));
  int P = 0;
  for (int u = 0; u < U; u++) {
    const char *pat = patterns + (size_t)u * n;
    int numFree = 0;
    for (int s = 0; s < n; s++) {
      if (!isDiploid[s]) { base[s] = pat[s]; continue; }
      genotypeBases(pat[s], base + s);
      if (orc_base_type(pat[s]) == 1 && !(breakSymmetries && breaks[(size_t)u * n + s])) freeSlot[numFree++] = s;
      s++;
    }
    const int phases = 1 << numFree;
    for (int k = 0; k < phases; k++) {
      char *col = phased + (size_t)(P + k) * n;
      memcpy(col, base, n);
      for (int j = 0; j < numFree; j++)
        if ((k >> j) & 1) { col[freeSlot[j]] = base[freeSlot[j] + 1]; col[freeSlot[j] + 1] = base[freeSlot[j]]; }
      numPhases[P + k] = k == 0 ? phases : 0;
    }
    P += phases;
  }
  free(base);
  free(freeSlot);
  free(breaks);
  return P;
}
