// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference header
//   /root/reference/gpytorch_lattice_kernel/cpp/permutohedral.h
// (found through the include path given by oracle/build_oracle.py; the header is
// compiled where it lies, nothing is copied) and exposes the intermediate lattice
// structure the reference never returns: per-point greedy/rank, per-vertex
// offsets/weights (the replay table), keys in first-touch order, the lattice
// values after splat and after blur, and the sliced output.
//
// Used by tests/golden/make_golden.py to produce the committed golden vectors
// and by tests (when oracle/_ref is built) to pin oracle/lattice_oracle.c bitwise.
#include <torch/extension.h>
#include <torch/torch.h>
#include <vector>
#include <cstring>

// The replay table and scale factors are private members (permutohedral.h:574-585);
// this translation unit is a test probe, so open them up rather than re-deriving them.
#define private public
#include "permutohedral.h"
#undef private

static at::Tensor ref_filter(at::Tensor src, at::Tensor ref, at::Tensor coeffs) {
    return PermutohedralLattice::filter(src, ref, coeffs);
}

// returns {greedy[N,d+1] i16, rank[N,d+1] i8, offsets[N,d+1] i32 (lattice index),
//          weights[N,d+1] f32, keys[M,d] i16, splatted[M,C] f32, blurred[M,C] f32,
//          out[N,C] f32, scale[d] f32}
static std::vector<at::Tensor> ref_structure(at::Tensor src, at::Tensor ref, at::Tensor coeffs) {
    src = src.contiguous();
    ref = ref.contiguous();
    coeffs = coeffs.contiguous();
    const int64_t n = src.size(0);
    const int vd = (int)src.size(1);
    const int d = (int)ref.size(1);
    TORCH_CHECK(ref.size(0) == n, "shape mismatch");
    const float *psrc = src.data_ptr<float>();
    const float *pref = ref.data_ptr<float>();

    auto greedy = torch::empty({n, d + 1}, torch::kInt16);
    auto rank = torch::empty({n, d + 1}, torch::kInt8);
    auto offsets = torch::empty({n, d + 1}, torch::kInt32);
    auto weights = torch::empty({n, d + 1}, torch::kFloat32);
    auto scale = torch::empty({d}, torch::kFloat32);
    auto out = torch::zeros({n, vd}, torch::kFloat32);

    PermutohedralLattice lat(d, vd, (int)n, coeffs);
    std::memcpy(scale.data_ptr<float>(), lat.scaleFactor, sizeof(float) * d);
    int16_t *pg = greedy.data_ptr<int16_t>();
    int8_t *pr = rank.data_ptr<int8_t>();
    for (int64_t i = 0; i < n; ++i) {
        lat.splat(const_cast<float *>(pref + i * d), const_cast<float *>(psrc + i * vd));
        for (int c = 0; c <= d; ++c) {
            pg[i * (d + 1) + c] = lat.greedy[c];
            pr[i * (d + 1) + c] = (int8_t)lat.rank[c];
        }
    }
    const int64_t M = lat.hashTable.size();
    int32_t *po = offsets.data_ptr<int32_t>();
    float *pw = weights.data_ptr<float>();
    for (int64_t i = 0; i < n * (d + 1); ++i) {
        po[i] = lat.replay[i].offset / vd;
        pw[i] = lat.replay[i].weight;
    }
    auto keys = torch::empty({M, d}, torch::kInt16);
    std::memcpy(keys.data_ptr<int16_t>(), lat.hashTable.getKeys(), sizeof(int16_t) * M * d);
    auto splatted = torch::empty({M, vd}, torch::kFloat32);
    std::memcpy(splatted.data_ptr<float>(), lat.hashTable.getValues(), sizeof(float) * M * vd);
    lat.blur(coeffs);
    auto blurred = torch::empty({M, vd}, torch::kFloat32);
    std::memcpy(blurred.data_ptr<float>(), lat.hashTable.getValues(), sizeof(float) * M * vd);
    lat.beginSlice();
    float *pout = out.data_ptr<float>();
    for (int64_t i = 0; i < n; ++i) lat.slice(pout + i * vd);
    return {greedy, rank, offsets, weights, keys, splatted, blurred, out, scale};
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("filter", &ref_filter, "reference PermutohedralLattice::filter (unmodified)");
    m.def("structure", &ref_structure, "reference lattice intermediates");
}
