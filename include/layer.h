// include/layer.h -- drop-in replacement for /root/reference/Sequential/layer.h on top of libpcnn.so.
//
// Same constants, same `class Layer` (members, constructor signature, setOutput / clear / bp_clear) and the same 18
// free functions with the same parameter types and order as the reference header [ref: Sequential/layer.h:12-414],
// so a driver written against the reference (Sequential/Main.cpp) compiles against this file unchanged apart from the
// two places where host code dereferences layer buffers (Main.cpp:168 vectorNorm(l_f.d_preact), Main.cpp:191
// res[i] = l_f.output[i]): as in the reference's own CUDA/ variant the seven Layer pointers are DEVICE pointers
// (cf. CUDA/layer_c.h:14-35) and those reads go through Layer::download().  Every function forwards to the C ABI in
// include/pcnn.h (cited per function); a non-zero status aborts with pcnn_last_error_string(), which keeps the
// reference's `void` signatures.
//
// Unlike the reference header this one has a working include guard and only inline definitions, so it may be
// included from several translation units.  Link with -lpcnn (parallel-cnn_b200/libpcnn.so).
#ifndef PCNN_DROPIN_LAYER_H_
#define PCNN_DROPIN_LAYER_H_

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "pcnn.h"

const static float dt = 1.0E-01f;          // [ref: layer.h:12]
const static float threshold = 1.0E-02f;   // [ref: layer.h:13]

namespace pcnn_dropin {
// One process-wide context, created on first use on the current device (the reference is global single-thread state).
inline pcnn_ctx *&ctx_slot() {
    static pcnn_ctx *c = nullptr;
    return c;
}
inline void check(int rc, const char *what) {
    if (rc != 0) {
        std::fprintf(stderr, "%s failed (%d): %s\n", what, rc, pcnn_last_error_string());
        std::abort();   // the reference's functions return void; there is nothing to return an error through
    }
}
inline pcnn_ctx *ctx() {
    pcnn_ctx *&c = ctx_slot();
    if (!c) check(pcnn_create(&c, -1, nullptr), "pcnn_create");
    return c;
}
}  // namespace pcnn_dropin

class Layer {   // [ref: layer.h:15-36]
public:
    int M, N, O;

    float *output;     // device pointers, zero-initialised (the reference's `new float[n]()`)
    float *preact;

    float *bias;
    float *weight;

    float *d_output;
    float *d_preact;
    float *d_weight;

    Layer(int M, int N, int O);
    ~Layer();

    void setOutput(float *data);
    void clear();
    void bp_clear();

    // additions: host access to device buffers (what CUDA/main.cu:220 does with cudaMemcpy)
    void download(float *host, const float *dev, int count) const {
        pcnn_dropin::check(pcnn_d2h(pcnn_dropin::ctx(), host, dev, sizeof(float) * (size_t)count), "pcnn_d2h");
    }
    void upload(float *dev, const float *host, int count) {
        pcnn_dropin::check(pcnn_h2d(pcnn_dropin::ctx(), dev, host, sizeof(float) * (size_t)count), "pcnn_h2d");
    }

private:
    Layer(const Layer &);              // the reference defines no copy; copying would double-free
    Layer &operator=(const Layer &);
};

// [ref: layer.h:39-55]  Buffers on the device; bias then M weights per neuron drawn from rand() on the HOST in the
// reference's order, so a driver with static Layer objects starts from the reference's seed-1 state.
inline Layer::Layer(int M, int N, int O) : M(M), N(N), O(O) {
    pcnn_ctx *c = pcnn_dropin::ctx();
    auto alloc = [&](float **p, int n) { pcnn_dropin::check(pcnn_malloc(c, (void **)p, sizeof(float) * (size_t)n), "pcnn_malloc"); };
    alloc(&output, O);
    alloc(&preact, O);
    alloc(&bias, N);
    alloc(&weight, M * N);
    alloc(&d_output, O);
    alloc(&d_preact, O);
    alloc(&d_weight, M * N);
    if (N > 0) {
        float *hb = new float[N];
        float *hw = new float[(size_t)M * N + 1];
        for (int i = 0; i < N; ++i) {
            hb[i] = 0.5f - static_cast<float>(rand()) / RAND_MAX;
            for (int j = 0; j < M; ++j) hw[i * M + j] = 0.5f - static_cast<float>(rand()) / RAND_MAX;
        }
        upload(bias, hb, N);
        if (M > 0) upload(weight, hw, M * N);
        delete[] hb;
        delete[] hw;
    }
}

inline Layer::~Layer() {   // [ref: layer.h:58-66]
    pcnn_ctx *c = pcnn_dropin::ctx_slot();
    if (!c) return;
    pcnn_free(c, output);
    pcnn_free(c, preact);
    pcnn_free(c, bias);
    pcnn_free(c, weight);
    pcnn_free(c, d_output);
    pcnn_free(c, d_preact);
    pcnn_free(c, d_weight);
}

inline void Layer::setOutput(float *data) {   // [ref: layer.h:68-70]; `data` is a HOST pointer as in Main.cpp:78
    pcnn_dropin::check(pcnn_h2d(pcnn_dropin::ctx(), output, data, sizeof(float) * (size_t)O), "pcnn_h2d");
}
inline void Layer::clear() {   // [ref: layer.h:72-75]
    pcnn_dropin::check(pcnn_memset0(pcnn_dropin::ctx(), output, sizeof(float) * (size_t)O), "pcnn_memset0");
    pcnn_dropin::check(pcnn_memset0(pcnn_dropin::ctx(), preact, sizeof(float) * (size_t)O), "pcnn_memset0");
}
inline void Layer::bp_clear() {   // [ref: layer.h:77-79]
    pcnn_dropin::check(pcnn_memset0(pcnn_dropin::ctx(), d_weight, sizeof(float) * (size_t)M * N), "pcnn_memset0");
}

// [ref: layer.h:81-83] scalar host helper, same expression as the reference (double exp, fp32 result)
inline float step_function(float v) { return 1 / (1 + exp(-v)); }

#define PCNN_FWD(call) pcnn_dropin::check((call), #call)

inline void apply_step_function(float *input, float *output, int N) {   // [ref: layer.h:85-89]
    PCNN_FWD(pcnn_apply_step_function(pcnn_dropin::ctx(), input, output, N));
}
inline void makeError(float *err, float *output, unsigned int Y, int N) {   // [ref: layer.h:91-95]
    PCNN_FWD(pcnn_make_error(pcnn_dropin::ctx(), err, output, Y, N));
}
inline void apply_grad(float *output, float *grad, int N) {   // [ref: layer.h:97-101]
    PCNN_FWD(pcnn_apply_grad(pcnn_dropin::ctx(), output, grad, N));
}
inline void fp_c1(const float input[28][28], float preact[6][24][24], const float weight[6][5][5], const float bias[6]) {   // [ref: layer.h:105]
    PCNN_FWD(pcnn_fp_c1(pcnn_dropin::ctx(), &input[0][0], &preact[0][0][0], &weight[0][0][0], bias, 1));
}
inline void fp_s1(const float input[6][24][24], float preact[6][6][6], const float weight[1][4][4], const float bias[1]) {   // [ref: layer.h:143]
    PCNN_FWD(pcnn_fp_s1(pcnn_dropin::ctx(), &input[0][0][0], &preact[0][0][0], &weight[0][0][0], bias, 1));
}
inline void fp_preact_f(const float input[6][6][6], float preact[10], const float weight[10][6][6][6]) {   // [ref: layer.h:184]
    PCNN_FWD(pcnn_fp_preact_f(pcnn_dropin::ctx(), &input[0][0][0], preact, &weight[0][0][0][0], 1));
}
inline void fp_bias_f(float preact[10], const float bias[10]) {   // [ref: layer.h:206]
    PCNN_FWD(pcnn_fp_bias_f(pcnn_dropin::ctx(), preact, bias, 1));
}
inline void bp_weight_f(float d_weight[10][6][6][6], const float d_preact[10], const float p_output[6][6][6]) {   // [ref: layer.h:214]
    PCNN_FWD(pcnn_bp_weight_f(pcnn_dropin::ctx(), &d_weight[0][0][0][0], d_preact, &p_output[0][0][0], 1));
}
inline void bp_bias_f(float bias[10], const float d_preact[10]) {   // [ref: layer.h:229]
    PCNN_FWD(pcnn_bp_bias_f(pcnn_dropin::ctx(), bias, d_preact, 1));
}
inline void bp_output_s1(float d_output[6][6][6], const float n_weight[10][6][6][6], const float nd_preact[10]) {   // [ref: layer.h:237]
    PCNN_FWD(pcnn_bp_output_s1(pcnn_dropin::ctx(), &d_output[0][0][0], &n_weight[0][0][0][0], nd_preact, 1));
}
inline void bp_preact_s1(float d_preact[6][6][6], const float d_output[6][6][6], const float preact[6][6][6]) {   // [ref: layer.h:260]
    PCNN_FWD(pcnn_bp_preact_s1(pcnn_dropin::ctx(), &d_preact[0][0][0], &d_output[0][0][0], &preact[0][0][0], 1));
}
inline void bp_weight_s1(float d_weight[1][4][4], const float d_preact[6][6][6], const float p_output[6][24][24]) {   // [ref: layer.h:272]
    PCNN_FWD(pcnn_bp_weight_s1(pcnn_dropin::ctx(), &d_weight[0][0][0], &d_preact[0][0][0], &p_output[0][0][0], 1));
}
inline void bp_bias_s1(float bias[1], const float d_preact[6][6][6]) {   // [ref: layer.h:302]
    PCNN_FWD(pcnn_bp_bias_s1(pcnn_dropin::ctx(), bias, &d_preact[0][0][0], 1));
}
inline void bp_output_c1(float d_output[6][24][24], const float n_weight[1][4][4], const float nd_preact[6][6][6]) {   // [ref: layer.h:319]
    PCNN_FWD(pcnn_bp_output_c1(pcnn_dropin::ctx(), &d_output[0][0][0], &n_weight[0][0][0], &nd_preact[0][0][0], 1));
}
inline void bp_preact_c1(float d_preact[6][24][24], const float d_output[6][24][24], const float preact[6][24][24]) {   // [ref: layer.h:348]
    PCNN_FWD(pcnn_bp_preact_c1(pcnn_dropin::ctx(), &d_preact[0][0][0], &d_output[0][0][0], &preact[0][0][0], 1));
}
inline void bp_weight_c1(float d_weight[6][5][5], const float d_preact[6][24][24], const float p_output[28][28]) {   // [ref: layer.h:371]
    PCNN_FWD(pcnn_bp_weight_c1(pcnn_dropin::ctx(), &d_weight[0][0][0], &d_preact[0][0][0], &p_output[0][0], 1));
}
inline void bp_bias_c1(float bias[6], const float d_preact[6][24][24]) {   // [ref: layer.h:398]
    PCNN_FWD(pcnn_bp_bias_c1(pcnn_dropin::ctx(), bias, &d_preact[0][0][0], 1));
}

#undef PCNN_FWD
#endif  // PCNN_DROPIN_LAYER_H_
