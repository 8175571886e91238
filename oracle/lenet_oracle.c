/*
 * oracle/lenet_oracle.c -- CPU restatement of the Parallel-CNN "Sequential" training path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA engine in
 * parallel-cnn_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load it.  Nothing in the product path links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks every function below bit-for-bit against
 *   (a) oracle/_ref/libref_seq.so = the unmodified reference sources compiled where they lie, and
 *   (b) tests/golden/ (npz files) generated from (a) by oracle/gen_golden.py, plus SURVEY.md Appendix B scalars.
 *
 * Every function cites the reference lines (under /root/reference/Sequential/) whose arithmetic and
 * evaluation ORDER it restates.  All storage is fp32; the sigmoid goes through double exp() exactly as the
 * reference's `1 / (1 + exp(-v))` does (layer.h:81-83, ::exp(double) overload).  Build with
 * `gcc -O2 -ffp-contract=off` and no -march/-ffast-math so that every multiply and add is rounded to
 * fp32 separately, as in the reference's own -O2 x86-64 build.
 *
 * Packed parameter / gradient vector used throughout the repo (2,343 floats):
 *     [   0, 150)  c1.weight [6][5][5]        (Main.cpp:18, layer.h:105)
 *     [ 150, 156)  c1.bias   [6]
 *     [ 156, 172)  s1.weight [1][4][4]        (Main.cpp:19, layer.h:143)
 *     [ 172, 173)  s1.bias   [1]
 *     [ 173,2333)  f.weight  [10][6][6][6]    (Main.cpp:20, layer.h:184)
 *     [2333,2343)  f.bias    [10]
 * The packed "gradient" g holds what the reference multiplies by dt: d_weight for the three weight blocks,
 * and for the bias blocks the RAW accumulators (c1: sum_xy d_preact, s1: sum d_preact, f: d_preact); the
 * per-block normalisation (/576, /216) is applied at update time in the reference's operand order
 * (layer.h:316, layer.h:412).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum {
    OFF_C1W = 0, OFF_C1B = 150, OFF_S1W = 156, OFF_S1B = 172, OFF_FW = 173, OFF_FB = 2333, NPARAM = 2343,
    N_IN = 784, N_C1 = 3456, N_S1 = 216, N_F = 10
};

/* activation record of one sample, in the order the engine's C-ABI also uses */
typedef struct {
    float c1_pre[N_C1], c1_out[N_C1];
    float s1_pre[N_S1], s1_out[N_S1];
    float f_pre[N_F], f_out[N_F];
} orc_acts;

/* backward record of one sample */
typedef struct {
    float f_dpre[N_F];
    float s1_dout[N_S1], s1_dpre[N_S1];
    float c1_dout[N_C1], c1_dpre[N_C1];
    float g[NPARAM];
    float err;
} orc_back;

int orc_sizeof_acts(void) { return (int)sizeof(orc_acts); }
int orc_sizeof_back(void) { return (int)sizeof(orc_back); }

/* layer.h:81-83 -- v is negated in fp32, exp and the quotient are double, result rounded to fp32 */
float orc_sigmoid(float v) {
    float nv = -v;
    return (float)(1.0 / (1.0 + exp((double)nv)));
}

/* layer.h:85-89 */
void orc_apply_step_function(const float *in, float *out, int n) {
    for (int t = 0; t < n; ++t) out[t] = orc_sigmoid(in[t]);
}

/* mnist.h:145 (u8 / 255.0 in double) followed by Main.cpp:64 (double -> float) */
void orc_u8_to_f32(const uint8_t *src, float *dst, long n) {
    for (long t = 0; t < n; ++t) {
        double d = src[t] / 255.0;
        dst[t] = (float)d;
    }
}

/* Layer::Layer, layer.h:39-55, for the three parametrised layers in construction order
 * (Main.cpp:17-20: l_input draws nothing because N == 0).  Each neuron draws its bias, then its M weights.
 * srand(1) reproduces the unseeded glibc sequence the static constructors see. */
static void draw_layer(float *w, float *b, int M, int N) {
    for (int n = 0; n < N; ++n) {
        b[n] = 0.5f - (float)rand() / RAND_MAX;
        for (int k = 0; k < M; ++k) w[n * M + k] = 0.5f - (float)rand() / RAND_MAX;
    }
}
void orc_init_params(float *p) {
    srand(1);
    draw_layer(p + OFF_C1W, p + OFF_C1B, 25, 6);
    draw_layer(p + OFF_S1W, p + OFF_S1B, 16, 1);
    draw_layer(p + OFF_FW, p + OFF_FB, 216, 10);
}

/* fp_c1, layer.h:105-140: valid 5x5 cross-correlation; private fp32 sum in (i,j) order, then bias. */
void orc_fp_c1(const float *in, float *pre, const float *w, const float *b) {
    for (int m = 0; m < 6; ++m)
        for (int x = 0; x < 24; ++x)
            for (int y = 0; y < 24; ++y) {
                float acc = 0.0f;
                for (int i = 0; i < 5; ++i)
                    for (int j = 0; j < 5; ++j) acc += in[(x + i) * 28 + (y + j)] * w[m * 25 + i * 5 + j];
                float v = 0.0f;
                v += acc;          /* preact zeroed then += sum (layer.h:108-130) */
                v += b[m];         /* bias pass (layer.h:133-139) */
                pre[m * 576 + x * 24 + y] = v;
            }
}

/* fp_s1, layer.h:143-181: shared 4x4 stride-4 weighted sum, weight first in the product. */
void orc_fp_s1(const float *in, float *pre, const float *w, const float *b) {
    for (int m = 0; m < 6; ++m)
        for (int x = 0; x < 6; ++x)
            for (int y = 0; y < 6; ++y) {
                float acc = 0.0f;
                for (int i = 0; i < 4; ++i)
                    for (int j = 0; j < 4; ++j) acc += w[i * 4 + j] * in[m * 576 + (x * 4 + i) * 24 + (y * 4 + j)];
                float v = 0.0f;
                v += acc;
                v += b[0];
                pre[m * 36 + x * 6 + y] = v;
            }
}

/* fp_preact_f, layer.h:184-203: accumulates straight into preact[o] over the flattened 216 inputs. */
void orc_fp_preact_f(const float *in, float *pre, const float *w) {
    for (int o = 0; o < 10; ++o) {
        float v = 0.0f;
        for (int k = 0; k < 216; ++k) v += w[o * 216 + k] * in[k];
        pre[o] = v;
    }
}

/* fp_bias_f, layer.h:206-211 */
void orc_fp_bias_f(float *pre, const float *b) {
    for (int o = 0; o < 10; ++o) pre[o] += b[o];
}

/* makeError, layer.h:91-95 */
void orc_make_error(float *err, const float *out, unsigned label, int n) {
    for (int t = 0; t < n; ++t) err[t] = ((unsigned)t == label) ? 1.0f - out[t] : -out[t];
}

/* vectorNorm, Main.cpp:28-34: fp32 sum of squares, sqrt through the double overload, fp32 result. */
float orc_vector_norm(const float *v, int n) {
    float s = 0.0f;
    for (int t = 0; t < n; ++t) s += v[t] * v[t];
    return (float)sqrt((double)s);
}

/* bp_weight_f, layer.h:214-227 (assignment, not accumulation) */
void orc_bp_weight_f(float *dw, const float *dpre, const float *pout) {
    for (int o = 0; o < 10; ++o)
        for (int k = 0; k < 216; ++k) dw[o * 216 + k] = dpre[o] * pout[k];
}

/* bp_bias_f, layer.h:229-234: in-place update with dt = 0.1f (layer.h:12) */
void orc_bp_bias_f(float *b, const float *dpre) {
    for (int o = 0; o < 10; ++o) b[o] += 1.0E-01f * dpre[o];
}

/* bp_output_s1, layer.h:237-257: output neuron o is the OUTER loop, so each d_output element
 * receives its ten terms in ascending o. */
void orc_bp_output_s1(float *dout, const float *nw, const float *ndpre) {
    for (int k = 0; k < 216; ++k) dout[k] = 0.0f;
    for (int o = 0; o < 10; ++o)
        for (int k = 0; k < 216; ++k) dout[k] += nw[o * 216 + k] * ndpre[o];
}

/* bp_preact_s1, layer.h:260-270: (d_output * o) * (1 - o), o recomputed from preact */
void orc_bp_preact_s1(float *dpre, const float *dout, const float *pre) {
    for (int k = 0; k < 216; ++k) {
        float o = orc_sigmoid(pre[k]);
        dpre[k] = dout[k] * o * (1 - o);
    }
}

/* bp_weight_s1, layer.h:272-300: un-normalised, accumulated in (m, x, y) order per tap */
void orc_bp_weight_s1(float *dw, const float *dpre, const float *pout) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float acc = 0.0f;
            for (int m = 0; m < 6; ++m)
                for (int x = 0; x < 6; ++x)
                    for (int y = 0; y < 6; ++y)
                        acc += dpre[m * 36 + x * 6 + y] * pout[m * 576 + (x * 4 + i) * 24 + (y * 4 + j)];
            dw[i * 4 + j] = acc;
        }
}

/* raw accumulator of bp_bias_s1, layer.h:303-314 */
float orc_bias_sum_s1(const float *dpre) {
    float s = 0.0f;
    for (int k = 0; k < 216; ++k) s += dpre[k];
    return s;
}
/* bp_bias_s1, layer.h:302-317: bias += dt * sum / 216 evaluated left to right in fp32 */
void orc_bp_bias_s1(float *b, const float *dpre) {
    float s = orc_bias_sum_s1(dpre);
    int total = 216;
    b[0] += 1.0E-01f * s / total;
}

/* bp_output_c1, layer.h:319-346: adjoint of the subsample; one term per element added to zero */
void orc_bp_output_c1(float *dout, const float *nw, const float *ndpre) {
    for (int t = 0; t < N_C1; ++t) dout[t] = 0.0f;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            for (int m = 0; m < 6; ++m)
                for (int x = 0; x < 6; ++x)
                    for (int y = 0; y < 6; ++y)
                        dout[m * 576 + (x * 4 + i) * 24 + (y * 4 + j)] += nw[i * 4 + j] * ndpre[m * 36 + x * 6 + y];
}

/* bp_preact_c1, layer.h:348-369: d_output * (s * (1 - s)), s rounded to fp32 first (layer.h:356) */
void orc_bp_preact_c1(float *dpre, const float *dout, const float *pre) {
    for (int t = 0; t < N_C1; ++t) {
        float s = orc_sigmoid(pre[t]);
        float ds = s * (1 - s);
        dpre[t] = dout[t] * ds;
    }
}

/* bp_weight_c1, layer.h:371-395: each product is divided by 576.0f before it is added */
void orc_bp_weight_c1(float *dw, const float *dpre, const float *pout) {
    const float d = 24.0f * 24.0f;
    for (int m = 0; m < 6; ++m)
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j) {
                float acc = 0.0f;
                for (int x = 0; x < 24; ++x)
                    for (int y = 0; y < 24; ++y)
                        acc += dpre[m * 576 + x * 24 + y] * pout[(x + i) * 28 + (y + j)] / d;
                dw[m * 25 + i * 5 + j] = acc;
            }
}

/* raw accumulators of bp_bias_c1, layer.h:400-410 */
void orc_bias_sum_c1(float *acc6, const float *dpre) {
    for (int m = 0; m < 6; ++m) {
        float s = 0.0f;
        for (int t = 0; t < 576; ++t) s += dpre[m * 576 + t];
        acc6[m] = s;
    }
}
/* bp_bias_c1, layer.h:398-414: bias += dt * acc / 576.0f */
void orc_bp_bias_c1(float *b, const float *dpre) {
    float acc[6];
    const float d = 24.0f * 24.0f;
    orc_bias_sum_c1(acc, dpre);
    for (int m = 0; m < 6; ++m) b[m] += 1.0E-01f * acc[m] / d;
}

/* apply_grad, layer.h:97-101 */
void orc_apply_grad(float *w, const float *g, int n) {
    for (int t = 0; t < n; ++t) w[t] += 1.0E-01f * g[t];
}

/* forward_pass, Main.cpp:59-105 (timers and clears omitted: every buffer is fully overwritten) */
void orc_forward(const float *p, const float *img, orc_acts *a) {
    orc_fp_c1(img, a->c1_pre, p + OFF_C1W, p + OFF_C1B);
    orc_apply_step_function(a->c1_pre, a->c1_out, N_C1);
    orc_fp_s1(a->c1_out, a->s1_pre, p + OFF_S1W, p + OFF_S1B);
    orc_apply_step_function(a->s1_pre, a->s1_out, N_S1);
    orc_fp_preact_f(a->s1_out, a->f_pre, p + OFF_FW);
    orc_fp_bias_f(a->f_pre, p + OFF_FB);
    orc_apply_step_function(a->f_pre, a->f_out, N_F);
}

/* learn() body for one sample with FROZEN parameters: Main.cpp:167-169 (makeError + vectorNorm) then the
 * bp_* chain of back_pass, Main.cpp:114-131, writing the packed gradient instead of touching p. */
void orc_backward(const float *p, const float *img, unsigned label, const orc_acts *a, orc_back *r) {
    orc_make_error(r->f_dpre, a->f_out, label, 10);
    r->err = orc_vector_norm(r->f_dpre, 10);
    orc_bp_weight_f(r->g + OFF_FW, r->f_dpre, a->s1_out);
    for (int o = 0; o < 10; ++o) r->g[OFF_FB + o] = r->f_dpre[o];
    orc_bp_output_s1(r->s1_dout, p + OFF_FW, r->f_dpre);
    orc_bp_preact_s1(r->s1_dpre, r->s1_dout, a->s1_pre);
    orc_bp_weight_s1(r->g + OFF_S1W, r->s1_dpre, a->c1_out);
    r->g[OFF_S1B] = orc_bias_sum_s1(r->s1_dpre);
    orc_bp_output_c1(r->c1_dout, p + OFF_S1W, r->s1_dpre);
    orc_bp_preact_c1(r->c1_dpre, r->c1_dout, a->c1_pre);
    orc_bp_weight_c1(r->g + OFF_C1W, r->c1_dpre, img);
    orc_bias_sum_c1(r->g + OFF_C1B, r->c1_dpre);
}

/* The update the reference performs spread over bp_bias_* (layer.h:229-234, 316, 412) and apply_grad
 * (Main.cpp:136-138), expressed on the packed vectors with step size lr (= dt at batch 1; dt / B for the
 * mini-batch extension).  Operand order follows the reference: (lr * g) / n. */
void orc_apply_update(float *p, const float *g, float lr) {
    for (int t = 0; t < 150; ++t) p[OFF_C1W + t] += lr * g[OFF_C1W + t];
    for (int t = 0; t < 6; ++t) p[OFF_C1B + t] += lr * g[OFF_C1B + t] / (24.0f * 24.0f);
    for (int t = 0; t < 16; ++t) p[OFF_S1W + t] += lr * g[OFF_S1W + t];
    p[OFF_S1B] += lr * g[OFF_S1B] / 216;
    for (int t = 0; t < 2160; ++t) p[OFF_FW + t] += lr * g[OFF_FW + t];
    for (int t = 0; t < 10; ++t) p[OFF_FB + t] += lr * g[OFF_FB + t];
}

/* One reference training step (learn() loop body, Main.cpp:157-171) on packed parameters; returns the
 * per-sample error norm that learn() sums. */
float orc_train_step(float *p, const float *img, unsigned label) {
    static orc_acts a;
    static orc_back r;
    orc_forward(p, img, &a);
    orc_backward(p, img, label, &a, &r);
    orc_apply_update(p, r.g, 1.0E-01f);
    return r.err;
}

/* learn(), Main.cpp:146-184, for n samples in dataset order starting from packed p; the fp32 running sum of
 * errors mirrors `err += tmp_err` (Main.cpp:169).  Returns err / n as learn() prints it. */
float orc_learn(float *p, const uint8_t *images_u8, const uint8_t *labels, long n) {
    float err = 0.0f;
    float img[N_IN];
    for (long s = 0; s < n; ++s) {
        orc_u8_to_f32(images_u8 + s * N_IN, img, N_IN);
        err += orc_train_step(p, img, labels[s]);
    }
    return n > 0 ? err / (float)n : 0.0f;   /* Main.cpp:173: float /= unsigned */
}

/* classify(), Main.cpp:186-200: strict '<' so the first maximum wins */
unsigned orc_classify(const float *p, const float *img) {
    static orc_acts a;
    orc_forward(p, img, &a);
    unsigned best = 0;
    for (unsigned t = 1; t < 10; ++t)
        if (a.f_out[best] < a.f_out[t]) best = t;
    return best;
}

/* test(), Main.cpp:202-214: number of misclassified samples */
long orc_test(const float *p, const uint8_t *images_u8, const uint8_t *labels, long n) {
    long wrong = 0;
    float img[N_IN];
    for (long s = 0; s < n; ++s) {
        orc_u8_to_f32(images_u8 + s * N_IN, img, N_IN);
        if (orc_classify(p, img) != labels[s]) ++wrong;
    }
    return wrong;
}

/* Mini-batch extension (not in the reference; SURVEY.md 8c "frozen-weight batch oracle"): per-sample packed
 * gradients computed exactly as above with the SAME parameters, accumulated in double.  err_sum receives the
 * double sum of the per-sample fp32 error norms. */
void orc_batch_grad(const float *p, const float *imgs, const uint8_t *labels, long B, double *g, double *err_sum) {
    static orc_acts a;
    static orc_back r;
    for (int t = 0; t < NPARAM; ++t) g[t] = 0.0;
    double es = 0.0;
    for (long s = 0; s < B; ++s) {
        orc_forward(p, imgs + s * N_IN, &a);
        orc_backward(p, imgs + s * N_IN, labels[s], &a, &r);
        for (int t = 0; t < NPARAM; ++t) g[t] += (double)r.g[t];
        es += (double)r.err;
    }
    if (err_sum) *err_sum = es;
}

/* ------------------------------------------------------------------------------------------------------
 * Extension ops named by BASELINE.json.north_star that the reference does not contain.
 * PARITY UNPINNED by the reference: these are self-written CPU definitions the CUDA kernels are tested
 * against.  Conventions are stated here and in DESIGN.md.
 * ---------------------------------------------------------------------------------------------------- */

/* x1: 2-D max-pool, window k, stride k, [C][H][W] -> [C][H/k][W/k], argmax cached as the flat index
 * (i * k + j) inside the window; ties resolve to the FIRST maximum in row-major (i, j) scan order, i.e.
 * strict '>' against the running maximum (the scan order of the PDF listing, section 4.3.2). */
void orc_maxpool_fwd(const float *in, float *out, int32_t *arg, int C, int H, int W, int k) {
    int Ho = H / k, Wo = W / k;
    for (int c = 0; c < C; ++c)
        for (int x = 0; x < Ho; ++x)
            for (int y = 0; y < Wo; ++y) {
                float best = in[c * H * W + (x * k) * W + (y * k)];
                int32_t bi = 0;
                for (int i = 0; i < k; ++i)
                    for (int j = 0; j < k; ++j) {
                        float v = in[c * H * W + (x * k + i) * W + (y * k + j)];
                        if (v > best) { best = v; bi = i * k + j; }
                    }
                out[c * Ho * Wo + x * Wo + y] = best;
                arg[c * Ho * Wo + x * Wo + y] = bi;
            }
}
/* x1 backward: routes each output gradient to its cached argmax position, zero elsewhere */
void orc_maxpool_bwd(const float *dout, const int32_t *arg, float *din, int C, int H, int W, int k) {
    int Ho = H / k, Wo = W / k;
    for (long t = 0; t < (long)C * H * W; ++t) din[t] = 0.0f;
    for (int c = 0; c < C; ++c)
        for (int x = 0; x < Ho; ++x)
            for (int y = 0; y < Wo; ++y) {
                int32_t a = arg[c * Ho * Wo + x * Wo + y];
                din[c * H * W + (x * k + a / k) * W + (y * k + a % k)] = dout[c * Ho * Wo + x * Wo + y];
            }
}

/* x2: softmax cross-entropy over n logits; probabilities and loss evaluated in double and rounded to fp32.
 * The returned d vector follows the reference's sign convention for d_preact (layer.h:93): onehot - p. */
float orc_softmax_ce(const float *z, unsigned label, int n, float *prob, float *d) {
    double mx = z[0];
    for (int t = 1; t < n; ++t) if (z[t] > mx) mx = z[t];
    double den = 0.0;
    for (int t = 0; t < n; ++t) den += exp((double)z[t] - mx);
    double loss = 0.0;
    for (int t = 0; t < n; ++t) {
        double pt = exp((double)z[t] - mx) / den;
        prob[t] = (float)pt;
        d[t] = (float)((((unsigned)t == label) ? 1.0 : 0.0) - pt);
        if ((unsigned)t == label) loss = -(((double)z[t] - mx) - log(den));
    }
    return (float)loss;
}

/* x3: generic direct convolution in the reference's conv semantics (valid, stride 1, cross-correlation,
 * layer.h:118-130 generalised to C input channels / K filters / RxS taps), NHWC activations, KRSC filters,
 * double accumulation.  x [N][H][W][C], w [K][R][S][C], y [N][P][Q][K] with P = H-R+1, Q = W-S+1. */
void orc_conv_fwd_nhwc(const float *x, const float *w, const float *bias, float *y,
                       int N, int H, int W, int C, int K, int R, int S) {
    int P = H - R + 1, Q = W - S + 1;
    for (int n = 0; n < N; ++n)
        for (int p = 0; p < P; ++p)
            for (int q = 0; q < Q; ++q)
                for (int k = 0; k < K; ++k) {
                    double acc = bias ? (double)bias[k] : 0.0;
                    for (int r = 0; r < R; ++r)
                        for (int s = 0; s < S; ++s)
                            for (int c = 0; c < C; ++c)
                                acc += (double)x[(((long)n * H + p + r) * W + q + s) * C + c] *
                                       (double)w[((k * R + r) * S + s) * C + c];
                    y[(((long)n * P + p) * Q + q) * K + k] = (float)acc;
                }
}
/* the same with a window step of `stride` in both directions (the reference's conv is stride 1):
 * y [N][(H-R)/stride+1][(W-S)/stride+1][K] */
void orc_conv_fwd_nhwc_strided(const float *x, const float *w, const float *bias, float *y,
                               int N, int H, int W, int C, int K, int R, int S, int stride) {
    int P = (H - R) / stride + 1, Q = (W - S) / stride + 1;
    for (int n = 0; n < N; ++n)
        for (int p = 0; p < P; ++p)
            for (int q = 0; q < Q; ++q)
                for (int k = 0; k < K; ++k) {
                    double acc = bias ? (double)bias[k] : 0.0;
                    for (int r = 0; r < R; ++r)
                        for (int s = 0; s < S; ++s)
                            for (int c = 0; c < C; ++c)
                                acc += (double)x[(((long)n * H + p * stride + r) * W + q * stride + s) * C + c] *
                                       (double)w[((k * R + r) * S + s) * C + c];
                    y[(((long)n * P + p) * Q + q) * K + k] = (float)acc;
                }
}
/* x3 wgrad: dw[k][r][s][c] = sum_{n,p,q} dy[n][p][q][k] * x[n][p+r][q+s][c]   (layer.h:371-395 without /576) */
void orc_conv_wgrad_nhwc(const float *x, const float *dy, float *dw,
                         int N, int H, int W, int C, int K, int R, int S) {
    int P = H - R + 1, Q = W - S + 1;
    for (int k = 0; k < K; ++k)
        for (int r = 0; r < R; ++r)
            for (int s = 0; s < S; ++s)
                for (int c = 0; c < C; ++c) {
                    double acc = 0.0;
                    for (int n = 0; n < N; ++n)
                        for (int p = 0; p < P; ++p)
                            for (int q = 0; q < Q; ++q)
                                acc += (double)dy[(((long)n * P + p) * Q + q) * K + k] *
                                       (double)x[(((long)n * H + p + r) * W + q + s) * C + c];
                    dw[((k * R + r) * S + s) * C + c] = (float)acc;
                }
}
/* x3 dgrad: dx[n][h][w][c] = sum_{k,r,s} dy[n][h-r][w-s][k] * w[k][r][s][c] over valid (h-r, w-s) */
void orc_conv_dgrad_nhwc(const float *dy, const float *w, float *dx,
                         int N, int H, int W, int C, int K, int R, int S) {
    int P = H - R + 1, Q = W - S + 1;
    for (int n = 0; n < N; ++n)
        for (int h = 0; h < H; ++h)
            for (int ww = 0; ww < W; ++ww)
                for (int c = 0; c < C; ++c) {
                    double acc = 0.0;
                    for (int r = 0; r < R; ++r) {
                        int p = h - r;
                        if (p < 0 || p >= P) continue;
                        for (int s = 0; s < S; ++s) {
                            int q = ww - s;
                            if (q < 0 || q >= Q) continue;
                            for (int k = 0; k < K; ++k)
                                acc += (double)dy[(((long)n * P + p) * Q + q) * K + k] *
                                       (double)w[((k * R + r) * S + s) * C + c];
                        }
                    }
                    dx[(((long)n * H + h) * W + ww) * C + c] = (float)acc;
                }
}
