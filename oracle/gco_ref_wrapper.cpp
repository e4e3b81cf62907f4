// ============================================================================
// gco_ref_wrapper.cpp — TEST-ONLY C wrapper around the REFERENCE's own
// alpha-expansion (GCO v3.0 as vendored by the reference under
// MultiH/MultiH/moduls/alpha_expansion/), compiled IN PLACE from
// /root/reference by oracle/Makefile into oracle/_ref/libgco_ref.so.
// No reference source is copied into this repository (GCO's licence forbids
// redistribution, GCoptimization.h:65-83); this file only contains our wrapper.
//
// It reproduces what MultiH::LabelingStep does with the library
// (MH.cpp:520-543): GCoptimizationGeneralGraph(N, L); dense site-major data
// costs (GCO.cpp:742 — same numbers the reference feeds through its dataEnergy
// callback, MH.cpp:522); Potts smoothness round(100*lambda) (MH.cpp:506-511);
// optional warm start (MH.cpp:525-529); setNeighbors(i, j) once per DIRECTED
// list entry (MH.cpp:532-540, so an unordered pair listed from both endpoints
// gets weight 2); expansion(iter, max_iter) (MH.cpp:543).
// ============================================================================
#include <cfloat>
#include <climits>
#include <cstring>
#include <cassert>
#include <cstdio>
#include <cstdint>
#include "GCoptimization.cpp"
#include "LinkedBlockList.cpp"

namespace {
struct PottsData { int w; };
int potts(int, int, int l1, int l2, void* d) { return l1 != l2 ? ((PottsData*)d)->w : 0; }
}

extern "C" int64_t gco_ref_expansion(int N, int L, const int* data_cost, int potts_weight, const int64_t* offsets,
                                     const int32_t* adj, const int32_t* init_labels /*or null*/, int max_iter,
                                     int32_t* labels_out, int* status) {
  *status = 0;
  int64_t energy = 0;
  // silence the reference's printf("cycle = ...") if any
  try {
    GCoptimizationGeneralGraph gc(N, L);
    gc.setDataCost(const_cast<int*>(data_cost));
    PottsData pd{potts_weight};
    gc.setSmoothCost(&potts, &pd);
    if (init_labels)
      for (int i = 0; i < N; ++i) gc.setLabel(i, init_labels[i]);
    for (int i = 0; i < N; ++i)
      for (int64_t e = offsets[i]; e < offsets[i + 1]; ++e)
        if (adj[e] != i) gc.setNeighbors(i, adj[e]);
    int iters = 0;
    energy = gc.expansion(iters, max_iter);
    for (int i = 0; i < N; ++i) labels_out[i] = gc.whatLabel(i);
  } catch (GCException& e) {
    *status = 1;
    std::fprintf(stderr, "gco_ref: %s\n", e.message);
  }
  return energy;
}
