// tree_ops.cuh — genealogy edit protocol shared by the host mirror and the device kernels.
//
// Semantics follow the reference's save/revert protocol (LocusDataLikelihood.c:768-1012, 1864-1906;
// restated in SURVEY.md Appendix A) but not its representation: instead of two node structs per node
// whose pointers are swapped, a genealogy is a pair of plain arrays (current / saved) and one flag
// byte per node:
//     bit 0  SEL     which of the node's two conditional-likelihood buffers is current
//     bit 1  RECALC  buffer was flipped in this proposal and must be (re)computed   (recalcConditionals[])
//     bit 2  SAVED   node fields were copied to the saved arrays in this proposal   (changedNodeIds[])
// "accept" clears RECALC/SAVED, "reject" restores SAVED nodes and flips RECALC buffers back — no
// conditional-likelihood data ever moves.  The same inline functions run on the host mirror (so the
// reference's getters stay plain host reads) and inside the device edit kernel (so only 24-byte edit
// records cross PCIe).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define GP_HD __host__ __device__ __forceinline__
#else
#define GP_HD inline
#endif

namespace gphocs {

enum : uint8_t { F_SEL = 1, F_RECALC = 2, F_SAVED = 4 };

enum OpType : int {
  OP_ADJUST_AGE = 0,  // a = node, x = new age            (adjustGenNodeAge,  .c:875)
  OP_SPR = 1,         // a = subtree root, b = target, x = age; status 0/1/2 (executeGenSPR, .c:931)
  OP_SCALE_ALL = 2,   // x = factor                       (scaleAllNodeAges,  .c:895, without the evaluation)
  OP_COMMIT = 3,      //                                   (resetSaved,        .c:852)
  OP_REVERT = 4,      //                                   (revertToSaved,     .c:768)
  OP_SET_RATE = 5,    // x = mutation rate                (setLocusMutationRate, .c:369)
};

struct Op {
  int locus;
  int type;
  int a, b;
  double x;
};

// One genealogy node: topology + flag byte in a single 8-byte record (one load per node on the device).
struct alignas(8) NodeRec {
  int16_t father, left, right;
  uint8_t flags;
  uint8_t pad;
};

// View over one locus' slice of the genealogy store.
struct TreeView {
  NodeRec *node;   // current [2n-1]
  NodeRec *saved;  // saved copies (their flag bytes are unused)
  double *age, *svAge;
  int *root, *savedRoot;
  double *lnL, *savedLnL, *rate;
  int numLeaves;
  int numPatterns;  // live phased patterns of this locus
};

// copyNodeConditionals (.c:1889-1906)
GP_HD void flipClv(const TreeView& t, int node) {
  if (t.numPatterns <= 0 || (t.node[node].flags & F_RECALC)) return;
  t.node[node].flags = (uint8_t)((t.node[node].flags ^ F_SEL) | F_RECALC);
}

// copyNodeToSaved (.c:1864-1876)
GP_HD void saveNode(const TreeView& t, int node, bool recalc) {
  if (recalc) flipClv(t, node);
  t.node[node].flags |= F_SAVED;
  t.svAge[node] = t.age[node];
  t.saved[node] = t.node[node];
}

GP_HD void adjustAge(const TreeView& t, int node, double age) {
  saveNode(t, node, true);
  t.age[node] = age;
}

GP_HD void scaleAll(const TreeView& t, double factor) {
  const int N = 2 * t.numLeaves - 1;
  for (int i = 0; i < N; i++) adjustAge(t, i, factor * t.age[i]);
}

// executeGenSPR (.c:931-1012)
GP_HD int spr(const TreeView& t, int sub, int target, double age) {
  NodeRec* nd = t.node;
  const int targetFather = nd[target].father;
  const int father = nd[sub].father;
  const int grandpa = nd[father].father;
  const int sibling = nd[father].left + nd[father].right - sub;
  adjustAge(t, father, age);
  if (target == sibling || target == father) return 0;
  saveNode(t, sibling, false);
  nd[sibling].father = (int16_t)grandpa;
  if (grandpa >= 0) {
    saveNode(t, grandpa, true);
    if (nd[grandpa].left == father) nd[grandpa].left = (int16_t)sibling;
    else nd[grandpa].right = (int16_t)sibling;
  }
  nd[father].father = (int16_t)targetFather;
  nd[father].left = (int16_t)sub;
  nd[father].right = (int16_t)target;
  if (target != grandpa) saveNode(t, target, false);
  nd[target].father = (int16_t)father;
  if (targetFather < 0) {
    *t.savedRoot = target;
    *t.root = father;
    return 1;
  }
  if (targetFather == sibling) flipClv(t, targetFather);
  else if (targetFather != grandpa) saveNode(t, targetFather, true);
  if (nd[targetFather].left == target) nd[targetFather].left = (int16_t)father;
  else nd[targetFather].right = (int16_t)father;
  if (grandpa < 0) {
    *t.savedRoot = father;
    *t.root = sibling;
    return 2;
  }
  return 0;
}

// resetSaved (.c:852-864), split into its per-node and per-locus parts so device code can spread the nodes over lanes
GP_HD void commitNode(const TreeView& t, int i) { t.node[i].flags &= F_SEL; }
GP_HD void commitLocus(const TreeView& t) {
  *t.savedRoot = -1;
  *t.savedLnL = *t.lnL;
}
GP_HD void commit(const TreeView& t) {
  const int N = 2 * t.numLeaves - 1;
  for (int i = 0; i < N; i++) commitNode(t, i);
  commitLocus(t);
}

// revertToSaved (.c:768-841).  The reference's copyAll branch (wholesale array swap after
// scaleAllNodeAges) is the same thing node by node because scaleAllNodeAges saved every node.
GP_HD void revertNode(const TreeView& t, int i) {
  NodeRec r = t.node[i];
  uint8_t f = r.flags;
  if (f & F_SAVED) {
    t.age[i] = t.svAge[i];
    r = t.saved[i];
  }
  if (f & F_RECALC) f ^= F_SEL;
  r.flags = f & F_SEL;
  t.node[i] = r;
}
GP_HD void revertLocus(const TreeView& t) {
  *t.lnL = *t.savedLnL;
  if (*t.savedRoot >= 0) {
    *t.root = *t.savedRoot;
    *t.savedRoot = -1;
  }
}
GP_HD void revert(const TreeView& t) {
  const int N = 2 * t.numLeaves - 1;
  revertLocus(t);
  for (int i = 0; i < N; i++) revertNode(t, i);
}

GP_HD int applyOp(const TreeView& t, const Op& op) {
  switch (op.type) {
    case OP_ADJUST_AGE: adjustAge(t, op.a, op.x); return 0;
    case OP_SPR: return spr(t, op.a, op.b, op.x);
    case OP_SCALE_ALL: scaleAll(t, op.x); return 0;
    case OP_COMMIT: commit(t); return 0;
    case OP_REVERT: revert(t); return 0;
    case OP_SET_RATE: *t.rate = op.x; return 0;
  }
  return -1;
}

}  // namespace gphocs
