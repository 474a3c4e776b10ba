// synth.cpp — synthetic workload generator (host only; measurement + test input infrastructure).
//
// Produces, for L independent loci under a population tree with optional migration bands:
//   * a genealogy per locus simulated by the structured coalescent with migration
//     (coalescence rate n(n-1)/theta per population, migration rate n*m per band, matching the
//     densities of the reference's gtreeLnLikelihood, patch.c:2702-2723),
//   * an alignment of S sites evolved under JC69 along that genealogy,
//   * the alignment compressed to weighted, JC-canonical site patterns, with every unphased
//     diploid heterozygote expanded into its 2^h phasings (the layout initializeLocusData consumes:
//     numPhases = 2^h on the first column of a phase group, 0 on the rest, one count per group —
//     LocusDataLikelihood.c:239-303; phase enumeration as AlignmentProcessor.c:2242-2339 but without
//     the optional symmetry breaking),
//   * optionally the sequence file in the reference's on-disk format (AlignmentProcessor.c:514-646)
//     so the reference's own ingest can be run on the same alignment,
//   * flattened per-population event chains of each genealogy (the snapshot the genealogy-likelihood
//     kernel consumes; ordering rule of createEvent, patch.c:1753-1802).
//
// Deterministic: every locus draws from its own generator seeded by (seed, locus).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int kMaxMigs = 10;  // reference cap per genealogy (patch.h:18)
constexpr double kOldAge = 999.0;  // patch.h:21

struct Rng {
  uint64_t s[4];
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
  Rng(uint64_t seed, uint64_t stream) {
    uint64_t x = seed * 0x2545F4914F6CDD1DULL + stream * 0x9E3779B97F4A7C15ULL + 1;
    for (int i = 0; i < 4; i++) s[i] = splitmix(x);
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  double uniform() { return (next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  double uniformPos() { double u; do u = uniform(); while (u <= 0.0); return u; }
  double expo(double rate) { return -std::log(uniformPos()) / rate; }
  int below(int n) { return (int)(uniform() * n); }
  double normal() {
    double u = uniformPos(), v = uniform();
    return std::sqrt(-2.0 * std::log(u)) * std::cos(6.283185307179586 * v);
  }
  double gamma(double shape) {  // Marsaglia-Tsang, shape >= 1 (boosted otherwise)
    if (shape < 1.0) return gamma(shape + 1.0) * std::pow(uniformPos(), 1.0 / shape);
    double d = shape - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
    for (;;) {
      double x = normal(), v = 1.0 + c * x;
      if (v <= 0) continue;
      v = v * v * v;
      double u = uniformPos();
      if (std::log(u) < 0.5 * x * x + d - d * v + d * std::log(v)) return d * v;
    }
  }
};

struct Model {
  int C, Q, B, S, diploid;
  std::vector<int> samplesPerPop, popFather, son0, son1, bandSrc, bandTgt;
  std::vector<double> popAge, sampleAge, theta, bandRate, bandStart, bandEnd;
  double rateShape, missingFrac;
  int n;  // haploid leaves
  int rootPop;
  std::vector<int> leafPop;
};

struct Mig { int branch, band, target, source; double age; };

struct Locus {
  std::vector<int> father, left, right, nodePop;
  std::vector<double> age;
  int root;
  double rate;
  std::vector<Mig> migs;
  // patterns
  int P = 0, U = 0;
  std::string chars;           // [P][n]
  std::vector<int> numPhases;  // [P]
  std::vector<int> counts;     // [U]
  std::vector<std::string> seqs;  // per sample (haploid: n, diploid: n/2) — only when kept
  // flattened events
  std::vector<int> popStart, evType, evId;
  std::vector<double> evTime;
};

struct Synth {
  Model m;
  std::vector<Locus> loci;
  bool keepSeqs;
};

enum { EV_COAL = 0, EV_IN_MIG, EV_OUT_MIG, EV_BAND_START, EV_BAND_END, EV_SAMPLES_START, EV_END_CHAIN };

const char kBase[4] = {'T', 'C', 'A', 'G'};

char iupac(int a, int b) {  // unordered pair of base indices (T=0,C=1,A=2,G=3) -> IUPAC symbol
  if (a == b) return kBase[a];
  if (a > b) std::swap(a, b);
  if (a == 0 && b == 1) return 'Y';
  if (a == 0 && b == 2) return 'W';
  if (a == 0 && b == 3) return 'K';
  if (a == 1 && b == 2) return 'M';
  if (a == 1 && b == 3) return 'S';
  return 'R';  // A/G
}

bool simulateGenealogy(const Model& m, Rng& rng, Locus& lc) {
  const int n = m.n, N = 2 * n - 1;
  lc.father.assign(N, -1); lc.left.assign(N, -1); lc.right.assign(N, -1);
  lc.nodePop.assign(N, -1); lc.age.assign(N, 0.0); lc.migs.clear();
  std::vector<std::vector<int>> lin(m.Q);
  std::vector<char> sampled(m.C, 0);
  for (int i = 0; i < n; i++) { lc.nodePop[i] = m.leafPop[i]; lc.age[i] = m.sampleAge[m.leafPop[i]]; }
  // breakpoints
  std::vector<double> bp;
  for (int p = m.C; p < m.Q; p++) bp.push_back(m.popAge[p]);
  for (int b = 0; b < m.B; b++) { bp.push_back(m.bandStart[b]); bp.push_back(m.bandEnd[b]); }
  for (int p = 0; p < m.C; p++) bp.push_back(m.sampleAge[p]);
  std::sort(bp.begin(), bp.end());
  bp.erase(std::unique(bp.begin(), bp.end()), bp.end());
  auto fatherAge = [&](int p) { return p == m.rootPop ? 1e300 : m.popAge[m.popFather[p]]; };
  auto applyBreak = [&](double t) {
    for (int p = 0; p < m.C; p++)
      if (!sampled[p] && m.sampleAge[p] <= t) {
        sampled[p] = 1;
        for (int i = 0; i < n; i++) if (m.leafPop[i] == p) lin[p].push_back(i);
      }
    // ancestral pops whose age has been reached absorb their sons (sons processed in id order;
    // pops are numbered so that ancestors of equal age still work by repeating until stable)
    bool moved = true;
    while (moved) {
      moved = false;
      for (int p = m.C; p < m.Q; p++)
        if (m.popAge[p] <= t)
          for (int s : {m.son0[p], m.son1[p]})
            if (!lin[s].empty() && (s >= m.C || sampled[s])) {
              lin[p].insert(lin[p].end(), lin[s].begin(), lin[s].end());
              lin[s].clear();
              moved = true;
            }
    }
  };
  double t = 0.0;
  size_t nextBp = 0;
  while (nextBp < bp.size() && bp[nextBp] <= t) nextBp++;
  applyBreak(t);
  int nextNode = n, remaining = n;
  std::vector<double> rates(m.Q + m.B);
  for (;;) {
    bool allSampled = true;
    for (int p = 0; p < m.C; p++) allSampled = allSampled && sampled[p];
    if (remaining == 1 && allSampled) break;
    double total = 0.0;
    for (int p = 0; p < m.Q; p++) {
      double r = 0.0;
      int k = (int)lin[p].size();
      if (k >= 2 && m.popAge[p] <= t && t < fatherAge(p)) r = k * (k - 1.0) / m.theta[p];
      rates[p] = r; total += r;
    }
    for (int b = 0; b < m.B; b++) {
      double r = 0.0;
      if (m.bandRate[b] > 0 && m.bandStart[b] <= t && t < m.bandEnd[b]) r = lin[m.bandTgt[b]].size() * m.bandRate[b];
      rates[m.Q + b] = r; total += r;
    }
    double tnext = nextBp < bp.size() ? bp[nextBp] : 1e300;
    double dt = total > 0 ? rng.expo(total) : 1e300;
    if (t + dt >= tnext) {
      if (tnext >= 1e299) return false;  // nothing can happen any more (should not occur)
      t = tnext; nextBp++;
      applyBreak(t);
      continue;
    }
    t += dt;
    double u = rng.uniform() * total;
    int which = 0;
    for (; which < m.Q + m.B - 1; which++) { if (u < rates[which]) break; u -= rates[which]; }
    while (rates[which] <= 0.0 && which > 0) which--;
    if (which < m.Q) {
      auto& v = lin[which];
      int i = rng.below((int)v.size()), j = rng.below((int)v.size() - 1);
      if (j >= i) j++;
      int a = v[i], b = v[j], id = nextNode++;
      lc.left[id] = a; lc.right[id] = b; lc.father[a] = id; lc.father[b] = id;
      lc.age[id] = t; lc.nodePop[id] = which;
      if (i < j) std::swap(i, j);
      v.erase(v.begin() + i); v.erase(v.begin() + j); v.push_back(id);
      remaining--;
    } else {
      int b = which - m.Q;
      auto& v = lin[m.bandTgt[b]];
      if ((int)lc.migs.size() >= kMaxMigs) return false;
      int i = rng.below((int)v.size());
      lc.migs.push_back({v[i], b, m.bandTgt[b], m.bandSrc[b], t});
      lin[m.bandSrc[b]].push_back(v[i]);
      v.erase(v.begin() + i);
    }
  }
  lc.root = nextNode - 1;
  return nextNode == N;
}

// canonical genotype column: labels 0..3 in order of first appearance, 4 = missing; for a diploid
// pair the two labels are stored sorted (phase is unknown)
void canonicalize(const Model& m, const std::vector<uint8_t>& col, const std::vector<char>& missing,
                  std::string& out) {
  int map[4] = {-1, -1, -1, -1}, nextLabel = 0;
  auto lab = [&](int b) { if (map[b] < 0) map[b] = nextLabel++; return map[b]; };
  out.assign(m.n, 0);
  for (int i = 0; i < m.n; i++) {
    if (missing[i]) { out[i] = 4; continue; }
    if (m.diploid && (i & 1) == 0) {
      int a = col[i], b = col[i + 1];
      if (a > b) std::swap(a, b);
      int la = lab(a), lb = lab(b);
      if (la > lb) std::swap(la, lb);
      out[i] = (char)la; out[i + 1] = (char)lb; i++;
    } else {
      out[i] = (char)lab(col[i]);
    }
  }
}

void simulateLocus(const Model& m, uint64_t seed, int locusIdx, Locus& lc, bool keepSeqs) {
  Rng rng(seed, (uint64_t)locusIdx);
  int attempts = 0;
  while (!simulateGenealogy(m, rng, lc)) { if (++attempts > 1000) { fprintf(stderr, "synth: genealogy simulation failed\n"); break; } }
  lc.rate = m.rateShape > 0 ? rng.gamma(m.rateShape) / m.rateShape : 1.0;
  const int n = m.n, N = 2 * n - 1, S = m.S;
  // per (locus, sample) missing data
  std::vector<char> missing(n, 0);
  if (m.missingFrac > 0) {
    int step = m.diploid ? 2 : 1;
    int nMissing = 0;
    for (int i = 0; i < n; i += step)
      if (rng.uniform() < m.missingFrac && nMissing + step < n - 1) {
        for (int k = 0; k < step; k++) missing[i + k] = 1;
        nMissing += step;
      }
  }
  // sequence evolution: states[node][site]
  std::vector<uint8_t> st((size_t)N * S);
  std::vector<int> varSites;
  for (int s = 0; s < S; s++) st[(size_t)lc.root * S + s] = (uint8_t)rng.below(4);
  // internal nodes were created in increasing age order: parents have larger ids than internal children
  std::vector<int> order;
  for (int i = N - 1; i >= 0; i--) order.push_back(i);
  for (int node : order) {
    if (node == lc.root) continue;
    int f = lc.father[node];
    uint8_t* dst = &st[(size_t)node * S];
    const uint8_t* src = &st[(size_t)f * S];
    memcpy(dst, src, S);
    double len = lc.rate * (lc.age[f] - lc.age[node]);
    double pChange = 0.75 * (1.0 - std::exp(-4.0 * len / 3.0));
    if (pChange <= 0) continue;
    double lg = std::log1p(-pChange);
    for (double pos = -1;;) {
      pos += 1.0 + std::floor(std::log(rng.uniformPos()) / lg);
      if (pos >= S) break;
      int s = (int)pos;
      dst[s] = (uint8_t)((dst[s] + 1 + rng.below(3)) & 3);
      varSites.push_back(s);
    }
  }
  std::sort(varSites.begin(), varSites.end());
  varSites.erase(std::unique(varSites.begin(), varSites.end()), varSites.end());
  // pattern compression at genotype level
  std::map<std::string, int> patt;
  std::vector<uint8_t> col(n);
  std::string key;
  for (int i = 0; i < n; i++) col[i] = 0;
  canonicalize(m, col, missing, key);
  patt[key] = S - (int)varSites.size();
  std::vector<std::string> orderKeys{key};
  for (int s : varSites) {
    for (int i = 0; i < n; i++) col[i] = st[(size_t)i * S + s];
    canonicalize(m, col, missing, key);
    auto it = patt.find(key);
    if (it == patt.end()) { patt[key] = 1; orderKeys.push_back(key); } else it->second++;
  }
  // phase expansion
  lc.P = 0; lc.U = 0; lc.chars.clear(); lc.numPhases.clear(); lc.counts.clear();
  for (auto& k : orderKeys) {
    int cnt = patt[k];
    if (cnt <= 0) continue;
    std::vector<int> hets;
    if (m.diploid)
      for (int i = 0; i + 1 < n; i += 2) if (k[i] != 4 && k[i] != k[i + 1]) hets.push_back(i);
    int h = (int)hets.size();
    if (h > 12) h = 12;  // cap at 4096 phasings (never reached at realistic theta)
    int phases = 1 << h;
    for (int mask = 0; mask < phases; mask++) {
      std::string ph = k;
      for (int j = 0; j < h; j++) if (mask >> j & 1) std::swap(ph[hets[j]], ph[hets[j] + 1]);
      for (int i = 0; i < n; i++) lc.chars.push_back(ph[i] == 4 ? 'N' : kBase[(int)ph[i]]);
      lc.numPhases.push_back(mask == 0 ? phases : 0);
      lc.P++;
    }
    lc.counts.push_back(cnt);
    lc.U++;
  }
  if (keepSeqs) {
    int step = m.diploid ? 2 : 1;
    lc.seqs.clear();
    for (int i = 0; i < n; i += step) {
      std::string s(S, 'N');
      if (!missing[i])
        for (int k = 0; k < S; k++) {
          int a = st[(size_t)i * S + k];
          s[k] = m.diploid ? iupac(a, st[(size_t)(i + 1) * S + k]) : kBase[a];
        }
      lc.seqs.push_back(s);
    }
  }
}

// Flattened event chains of one genealogy. Within a population events are ordered by age; events
// created later by constructEventChain land BEFORE earlier-created ones of equal age
// (createEvent walks while elapsed_time < remaining, patch.c:1778-1785), creation order being:
// band start/end, samples start, migrations, coalescences (patch.c:1996-2120).
void buildEvents(const Model& m, Locus& lc) {
  struct Ev { double age; int seq; int type; int id; };
  std::vector<std::vector<Ev>> per(m.Q);
  int seq = 0;
  for (int b = 0; b < m.B; b++) {
    per[m.bandTgt[b]].push_back({m.bandStart[b], seq++, EV_BAND_START, b});
    per[m.bandTgt[b]].push_back({m.bandEnd[b], seq++, EV_BAND_END, b});
  }
  for (int p = 0; p < m.C; p++) per[p].push_back({m.sampleAge[p], seq++, EV_SAMPLES_START, p});
  for (auto& g : lc.migs) {
    per[g.target].push_back({g.age, seq++, EV_IN_MIG, g.band});
    per[g.source].push_back({g.age, seq++, EV_OUT_MIG, g.band});
  }
  for (int node = m.n; node < 2 * m.n - 1; node++) per[lc.nodePop[node]].push_back({lc.age[node], seq++, EV_COAL, node});
  lc.popStart.assign(m.Q + 1, 0); lc.evType.clear(); lc.evId.clear(); lc.evTime.clear();
  for (int p = 0; p < m.Q; p++) {
    lc.popStart[p] = (int)lc.evType.size();
    auto& v = per[p];
    std::stable_sort(v.begin(), v.end(), [](const Ev& a, const Ev& b) { return a.age < b.age || (a.age == b.age && a.seq > b.seq); });
    double prev = m.popAge[p];
    for (auto& e : v) {
      lc.evType.push_back(e.type); lc.evId.push_back(e.id); lc.evTime.push_back(e.age - prev);
      prev = e.age;
    }
    double end = (p == m.rootPop) ? kOldAge : m.popAge[m.popFather[p]];
    lc.evType.push_back(EV_END_CHAIN); lc.evId.push_back(p); lc.evTime.push_back(end - prev);
  }
  lc.popStart[m.Q] = (int)lc.evType.size();
}

}  // namespace

extern "C" {

// popAge[Q]: tau of ancestral pops (0 for current pops); sampleAge[C]; samplesPerPop[C] counts haploid leaves.
void* synth_create(int numCurPops, int numBands, int numSites, int diploid, const int* samplesPerPop,
                   const int* popFather, const int* popSon0, const int* popSon1, const double* popAge,
                   const double* sampleAge, const double* theta, const int* bandSource, const int* bandTarget,
                   const double* bandRate, double rateShape, double missingFrac, int numLoci, uint64_t seed,
                   int keepSeqs, int nthreads) {
  Synth* h = new Synth();
  Model& m = h->m;
  m.C = numCurPops; m.Q = 2 * numCurPops - 1; m.B = numBands; m.S = numSites; m.diploid = diploid;
  m.samplesPerPop.assign(samplesPerPop, samplesPerPop + m.C);
  m.popFather.assign(popFather, popFather + m.Q);
  m.son0.assign(popSon0, popSon0 + m.Q); m.son1.assign(popSon1, popSon1 + m.Q);
  m.popAge.assign(popAge, popAge + m.Q);
  m.sampleAge.assign(sampleAge, sampleAge + m.C);
  m.theta.assign(theta, theta + m.Q);
  if (m.B > 0) {
    m.bandSrc.assign(bandSource, bandSource + m.B); m.bandTgt.assign(bandTarget, bandTarget + m.B);
    m.bandRate.assign(bandRate, bandRate + m.B);
  }
  m.rateShape = rateShape; m.missingFrac = missingFrac;
  m.rootPop = -1;
  for (int p = 0; p < m.Q; p++) if (m.popFather[p] < 0) m.rootPop = p;
  // current pops start at age 0 and carry a separate sample age (MCMCcontrol.c:850,894)
  for (int p = 0; p < m.C; p++) m.popAge[p] = 0.0;
  m.bandStart.resize(m.B); m.bandEnd.resize(m.B);
  for (int b = 0; b < m.B; b++) {  // PopulationTree.c:439-462
    int s = m.bandSrc[b], t = m.bandTgt[b];
    m.bandStart[b] = std::max(m.popAge[s], m.popAge[t]);
    m.bandEnd[b] = std::min(m.popAge[m.popFather[s]], m.popAge[m.popFather[t]]);
    if (m.bandStart[b] >= m.bandEnd[b]) m.bandStart[b] = m.bandEnd[b] = m.popAge[t];
  }
  m.n = 0;
  for (int p = 0; p < m.C; p++) for (int k = 0; k < m.samplesPerPop[p]; k++) { m.leafPop.push_back(p); m.n++; }
  h->keepSeqs = keepSeqs != 0;
  h->loci.resize(numLoci);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 64)
  for (int l = 0; l < numLoci; l++) {
    simulateLocus(m, seed, l, h->loci[l], h->keepSeqs);
    buildEvents(m, h->loci[l]);
  }
  return h;
}

void synth_free(void* hh) { delete (Synth*)hh; }
int synth_num_leaves(void* hh) { return ((Synth*)hh)->m.n; }

// totals: [0]=sum P, [1]=sum U, [2]=sum events, [3]=sum migs
void synth_totals(void* hh, long long* out) {
  Synth* h = (Synth*)hh;
  long long p = 0, u = 0, e = 0, g = 0;
  for (auto& lc : h->loci) { p += lc.P; u += lc.U; e += (long long)lc.evType.size(); g += (long long)lc.migs.size(); }
  out[0] = p; out[1] = u; out[2] = e; out[3] = g;
}

// CSR export. pattStart[L+1], unphStart[L+1]; chars[sumP*n]; numPhases[sumP]; counts[sumU];
// trees: father/left/right/nodePop [L*N] int32, age [L*N], root[L], rate[L].
void synth_export(void* hh, long long* pattStart, long long* unphStart, char* chars, int* numPhases, int* counts,
                  int* father, int* left, int* right, int* nodePop, double* age, int* root, double* rate) {
  Synth* h = (Synth*)hh;
  const int n = h->m.n, N = 2 * n - 1;
  long long p = 0, u = 0;
  for (size_t l = 0; l < h->loci.size(); l++) {
    Locus& lc = h->loci[l];
    pattStart[l] = p; unphStart[l] = u;
    memcpy(chars + p * n, lc.chars.data(), (size_t)lc.P * n);
    memcpy(numPhases + p, lc.numPhases.data(), sizeof(int) * lc.P);
    memcpy(counts + u, lc.counts.data(), sizeof(int) * lc.U);
    p += lc.P; u += lc.U;
    for (int i = 0; i < N; i++) {
      father[l * N + i] = lc.father[i]; left[l * N + i] = lc.left[i]; right[l * N + i] = lc.right[i];
      nodePop[l * N + i] = lc.nodePop[i]; age[l * N + i] = lc.age[i];
    }
    root[l] = lc.root; rate[l] = lc.rate;
  }
  pattStart[h->loci.size()] = p; unphStart[h->loci.size()] = u;
}

// events: evStart[L+1]; popStart[L*(Q+1)] (offsets relative to the locus' first event);
// evType/evId [sumE] int32; evTime[sumE].  migs: migStart[L+1]; branch/band/target/source, age.
void synth_export_events(void* hh, long long* evStart, int* popStart, int* evType, int* evId, double* evTime,
                         long long* migStart, int* migBranch, int* migBand, int* migTarget, int* migSource,
                         double* migAge) {
  Synth* h = (Synth*)hh;
  const int Q = h->m.Q;
  long long e = 0, g = 0;
  for (size_t l = 0; l < h->loci.size(); l++) {
    Locus& lc = h->loci[l];
    evStart[l] = e; migStart[l] = g;
    for (int p = 0; p <= Q; p++) popStart[l * (Q + 1) + p] = lc.popStart[p];
    for (size_t k = 0; k < lc.evType.size(); k++) { evType[e] = lc.evType[k]; evId[e] = lc.evId[k]; evTime[e] = lc.evTime[k]; e++; }
    for (auto& mg : lc.migs) { migBranch[g] = mg.branch; migBand[g] = mg.band; migTarget[g] = mg.target; migSource[g] = mg.source; migAge[g] = mg.age; g++; }
  }
  evStart[h->loci.size()] = e; migStart[h->loci.size()] = g;
}

void synth_band_times(void* hh, double* start, double* end) {
  Synth* h = (Synth*)hh;
  for (int b = 0; b < h->m.B; b++) { start[b] = h->m.bandStart[b]; end[b] = h->m.bandEnd[b]; }
}

// sequence file in the reference's format; sampleNames = numSamples C strings in leaf order
int synth_write_seqfile(void* hh, const char* path, const char* const* sampleNames) {
  Synth* h = (Synth*)hh;
  if (!h->keepSeqs) return -1;
  FILE* f = fopen(path, "w");
  if (!f) return -2;
  fprintf(f, "%d\n\n", (int)h->loci.size());
  for (size_t l = 0; l < h->loci.size(); l++) {
    Locus& lc = h->loci[l];
    fprintf(f, "locus%d %d %d\n", (int)l + 1, (int)lc.seqs.size(), h->m.S);
    for (size_t s = 0; s < lc.seqs.size(); s++) fprintf(f, "%s\t%s\n", sampleNames[s], lc.seqs[s].c_str());
    fprintf(f, "\n");
  }
  fclose(f);
  return 0;
}

}  // extern "C"
