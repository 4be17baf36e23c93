// TEST INFRASTRUCTURE ONLY — brute-force CPU restatement of simple-knn's distCUDA2
// (gaussian_splatting/submodules/simple-knn/simple_knn.cu:131-183,185-221; spatial.cu:15-26).
//
// The reference finds, exactly, the three smallest squared distances from each point to the other points
// (updateKBest<3>, :131-146, candidates i != idx) and stores (best[0]+best[1]+best[2])/3.0f (:183).  Its Morton
// order and box walk only decide WHICH pairs are evaluated, never the value, so an all-pairs scan with the same
// per-pair arithmetic gives the same bits.  Per-pair arithmetic as the reference build for sm_100a contracts it
// (oracle/_ref/simple_knn.sass, boxMeanDist): d = c - q per axis; dist = fma(dz, dz, fma(dx, dx, dy*dy)).
// Pinned on the GPU box against the reference build itself (tests/test_knn.py).  Never linked into the product.
#include <cfloat>
#include <cmath>
#include <cstdint>

extern "C" void knn_oracle_dist2(const float* pts, int64_t P, float* out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < P; i++) {
    const float qx = pts[3 * i], qy = pts[3 * i + 1], qz = pts[3 * i + 2];
    float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;
    for (int64_t j = 0; j < P; j++) {
      if (j == i) continue;
      const float dx = pts[3 * j] - qx, dy = pts[3 * j + 1] - qy, dz = pts[3 * j + 2] - qz;
      float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
      if (b0 > d) { float t = b0; b0 = d; d = t; }
      if (b1 > d) { float t = b1; b1 = d; d = t; }
      if (b2 > d) { b2 = d; }
    }
    out[i] = ((b0 + b1) + b2) / 3.0f;
  }
}
