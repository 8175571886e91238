// oracle/ref_harness.cpp -- C entry points around the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY (see oracle/lenet_oracle.c header).  This translation unit contains no
// reference code: it #includes /root/reference/Sequential/Main.cpp (which in turn includes layer.h and
// mnist.h) from where it lies, via -I, and is compiled by oracle/Makefile into the git-ignored
// oracle/_ref/libref_seq.so.  `main` is renamed so the driver's static functions (forward_pass,
// back_pass, learn, test, classify) and its static Layer objects become callable from this TU.
// layer.h has a broken include guard and non-inline definitions, so it must be
// included exactly once -- which Main.cpp already does.
//
// Build flags are the reference's "parity flags": g++ -O2, no -march, no -ffast-math.
#define main pcnn_reference_main_unused
#include "Main.cpp"
#undef main

#include <cstdint>
#include <chrono>

namespace {
// Repeats the draw order of the Layer constructor for the layers already constructed (the loop of
// Layer::Layer applied to existing storage) so tests can return to the seed-1 state at any time.
void redraw(Layer &l) {
    for (int n = 0; n < l.N; ++n) {
        l.bias[n] = 0.5f - static_cast<float>(rand()) / RAND_MAX;
        for (int k = 0; k < l.M; ++k) l.weight[n * l.M + k] = 0.5f - static_cast<float>(rand()) / RAND_MAX;
    }
}
mnist_data *staged = nullptr;
long staged_cap = 0;
mnist_data *stage_u8(const uint8_t *images, const uint8_t *labels, long n) {
    if (n > staged_cap) {
        free(staged);
        staged = (mnist_data *)malloc(sizeof(mnist_data) * (size_t)n);
        staged_cap = n;
    }
    for (long s = 0; s < n; ++s) {
        for (int t = 0; t < 784; ++t) staged[s].data[t / 28][t % 28] = images[s * 784 + t] / 255.0;  // as mnist_load
        staged[s].label = labels[s];
    }
    return staged;
}
}  // namespace

extern "C" {

int ref_sizeof_mnist_data() { return (int)sizeof(mnist_data); }

// packed order used across the repo: c1.w | c1.b | s1.w | s1.b | f.w | f.b
void ref_get_params(float *p) {
    memcpy(p, l_c1.weight, 150 * 4);        memcpy(p + 150, l_c1.bias, 6 * 4);
    memcpy(p + 156, l_s1.weight, 16 * 4);   memcpy(p + 172, l_s1.bias, 1 * 4);
    memcpy(p + 173, l_f.weight, 2160 * 4);  memcpy(p + 2333, l_f.bias, 10 * 4);
}
void ref_set_params(const float *p) {
    memcpy(l_c1.weight, p, 150 * 4);        memcpy(l_c1.bias, p + 150, 6 * 4);
    memcpy(l_s1.weight, p + 156, 16 * 4);   memcpy(l_s1.bias, p + 172, 1 * 4);
    memcpy(l_f.weight, p + 173, 2160 * 4);  memcpy(l_f.bias, p + 2333, 10 * 4);
}
void ref_reset_params() {
    srand(1);
    redraw(l_c1);
    redraw(l_s1);
    redraw(l_f);
}

// forward_pass on one image given as the loader would hold it (double, u8/255.0)
void ref_forward_u8(const uint8_t *img) {
    mnist_data *d = stage_u8(img, img /*label unused*/, 1);
    forward_pass(d[0].data);
}
// c1.preact, c1.output, s1.preact, s1.output, f.preact, f.output (7,364 floats)
void ref_get_acts(float *a) {
    memcpy(a, l_c1.preact, 3456 * 4);          memcpy(a + 3456, l_c1.output, 3456 * 4);
    memcpy(a + 6912, l_s1.preact, 216 * 4);    memcpy(a + 7128, l_s1.output, 216 * 4);
    memcpy(a + 7344, l_f.preact, 10 * 4);      memcpy(a + 7354, l_f.output, 10 * 4);
}
// f.d_preact, s1.d_output, s1.d_preact, c1.d_output, c1.d_preact, c1.d_weight, s1.d_weight, f.d_weight
void ref_get_back(float *b) {
    memcpy(b, l_f.d_preact, 10 * 4);
    memcpy(b + 10, l_s1.d_output, 216 * 4);       memcpy(b + 226, l_s1.d_preact, 216 * 4);
    memcpy(b + 442, l_c1.d_output, 3456 * 4);     memcpy(b + 3898, l_c1.d_preact, 3456 * 4);
    memcpy(b + 7354, l_c1.d_weight, 150 * 4);     memcpy(b + 7504, l_s1.d_weight, 16 * 4);
    memcpy(b + 7520, l_f.d_weight, 2160 * 4);
}

// The body of learn()'s per-sample loop, made of the driver's own functions in the driver's order.
static float one_step(mnist_data *d) {
    forward_pass(d->data);
    l_f.bp_clear();
    l_s1.bp_clear();
    l_c1.bp_clear();
    makeError(l_f.d_preact, l_f.output, d->label, 10);
    float e = vectorNorm(l_f.d_preact, 10);
    back_pass();
    return e;
}
float ref_train_step_u8(const uint8_t *img, unsigned label) {
    uint8_t lab = (uint8_t)label;
    mnist_data *d = stage_u8(img, &lab, 1);
    return one_step(d);
}
// n samples in order; returns err / n the way learn() computes it (fp32 running sum); seconds_out = wall time
float ref_learn_loop_u8(const uint8_t *images, const uint8_t *labels, long n, double *seconds_out) {
    mnist_data *d = stage_u8(images, labels, n);
    auto t0 = std::chrono::steady_clock::now();
    float err = 0.0f;
    for (long s = 0; s < n; ++s) err += one_step(&d[s]);
    auto t1 = std::chrono::steady_clock::now();
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    unsigned cnt = (unsigned)n;
    err /= cnt;
    return err;
}
// The driver's own learn() (prints "Learning", "error: ...", " Time - ...") on a staged set.
void ref_learn_driver_u8(const uint8_t *images, const uint8_t *labels, long n) {
    train_set = stage_u8(images, labels, n);
    train_cnt = (unsigned)n;
    learn();
    fflush(stdout);
}
// classify() over n samples -> number of mismatches (what test() counts)
long ref_test_u8(const uint8_t *images, const uint8_t *labels, long n) {
    mnist_data *d = stage_u8(images, labels, n);
    long wrong = 0;
    for (long s = 0; s < n; ++s)
        if (classify(d[s].data) != d[s].label) ++wrong;
    return wrong;
}
unsigned ref_classify_u8(const uint8_t *img) {
    mnist_data *d = stage_u8(img, img, 1);
    return classify(d[0].data);
}

// ---- op-level pass-throughs (flat pointers cast exactly as Main.cpp casts them, Main.cpp:81-131) ----
float ref_step_function(float v) { return step_function(v); }
void ref_apply_step_function(float *in, float *out, int n) { apply_step_function(in, out, n); }
void ref_makeError(float *err, float *out, unsigned y, int n) { makeError(err, out, y, n); }
void ref_apply_grad(float *w, float *g, int n) { apply_grad(w, g, n); }
float ref_vectorNorm(float *v, int n) { return vectorNorm(v, n); }
void ref_fp_c1(const float *in, float *pre, const float *w, const float *b) {
    fp_c1((const float(*)[28])in, (float(*)[24][24])pre, (const float(*)[5][5])w, b);
}
void ref_fp_s1(const float *in, float *pre, const float *w, const float *b) {
    fp_s1((const float(*)[24][24])in, (float(*)[6][6])pre, (const float(*)[4][4])w, b);
}
void ref_fp_preact_f(const float *in, float *pre, const float *w) {
    fp_preact_f((const float(*)[6][6])in, pre, (const float(*)[6][6][6])w);
}
void ref_fp_bias_f(float *pre, const float *b) { fp_bias_f(pre, b); }
void ref_bp_weight_f(float *dw, const float *dpre, const float *pout) {
    bp_weight_f((float(*)[6][6][6])dw, dpre, (const float(*)[6][6])pout);
}
void ref_bp_bias_f(float *b, const float *dpre) { bp_bias_f(b, dpre); }
void ref_bp_output_s1(float *dout, const float *nw, const float *ndpre) {
    bp_output_s1((float(*)[6][6])dout, (const float(*)[6][6][6])nw, ndpre);
}
void ref_bp_preact_s1(float *dpre, const float *dout, const float *pre) {
    bp_preact_s1((float(*)[6][6])dpre, (const float(*)[6][6])dout, (const float(*)[6][6])pre);
}
void ref_bp_weight_s1(float *dw, const float *dpre, const float *pout) {
    bp_weight_s1((float(*)[4][4])dw, (const float(*)[6][6])dpre, (const float(*)[24][24])pout);
}
void ref_bp_bias_s1(float *b, const float *dpre) { bp_bias_s1(b, (const float(*)[6][6])dpre); }
void ref_bp_output_c1(float *dout, const float *nw, const float *ndpre) {
    bp_output_c1((float(*)[24][24])dout, (const float(*)[4][4])nw, (const float(*)[6][6])ndpre);
}
void ref_bp_preact_c1(float *dpre, const float *dout, const float *pre) {
    bp_preact_c1((float(*)[24][24])dpre, (const float(*)[24][24])dout, (const float(*)[24][24])pre);
}
void ref_bp_weight_c1(float *dw, const float *dpre, const float *pout) {
    bp_weight_c1((float(*)[5][5])dw, (const float(*)[24][24])dpre, (const float(*)[28])pout);
}
void ref_bp_bias_c1(float *b, const float *dpre) { bp_bias_c1(b, (const float(*)[24][24])dpre); }

}  // extern "C"
