/*
 * oracle/lenet5_oracle.c -- CPU definition of the LeNet-5-style variant ("a second conv layer, true LeNet-5
 * C3/S4", which the reference does not have).
 *
 * TEST INFRASTRUCTURE ONLY, and PARITY UNPINNED: the reference contains no second convolution, no 2x2 subsample and no
 * convolution input gradient, so nothing in /root/reference can pin this file.  It generalises the reference's own layer
 * functions BY RULE (each rule cites the reference lines it extends) and is the checker of csrc/lenet5_kernels.cu.
 *
 * Network (28x28 input, all activations sigmoid, loss as the reference: d_preact_f = onehot - output, layer.h:91-95):
 *     c1: 6 maps 5x5 valid            -> [6][24][24]    (layer.h:105-140, unchanged)
 *     s2: shared 2x2/2 weighted sum   -> [6][12][12]    (rule of fp_s1, layer.h:143-181, window 2 instead of 4)
 *     c3: 16 maps 5x5 over 6 channels -> [16][8][8]     (rule of fp_c1 with a channel sum, order c, i, j; bias last)
 *     s4: shared 2x2/2 weighted sum   -> [16][4][4]
 *     f : 256 -> 10                                     (layer.h:184-211)
 * Packed parameters (5,152 floats): c1w 150 | c1b 6 | s2w 4 | s2b 1 | c3w [16][6][5][5] 2400 | c3b 16 | s4w 4 | s4b 1 |
 * fw [10][256] 2560 | fb 10.  The packed gradient holds what the reference multiplies by dt (the NEGATIVE gradient) with
 * the reference's normalisation rules extended: conv weight terms divided by the map size as they are added (layer.h:389:
 * /576 for c1, /64 for c3), conv and subsample bias blocks hold RAW sums and are divided at update time by the map size
 * (layer.h:412) / the layer's output count (layer.h:316), subsample and f weights un-normalised (layer.h:293, 222).
 * The convolution input gradient (s2's d_output) has no counterpart in the reference; it is the plain adjoint of c3.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

enum { L5_C1W = 0, L5_C1B = 150, L5_S2W = 156, L5_S2B = 160, L5_C3W = 161, L5_C3B = 2561, L5_S4W = 2577, L5_S4B = 2581,
       L5_FW = 2582, L5_FB = 5142, L5_NPARAM = 5152 };

typedef struct {
    float c1_pre[6][24][24], c1_out[6][24][24], s2_pre[6][12][12], s2_out[6][12][12];
    float c3_pre[16][8][8], c3_out[16][8][8], s4_pre[16][4][4], s4_out[16][4][4], f_pre[10], f_out[10];
} l5_acts;

int l5_nparam(void) { return L5_NPARAM; }
int l5_sizeof_acts(void) { return (int)sizeof(l5_acts); }

static float sigm(float v) { return (float)(1.0 / (1.0 + exp((double)(-v)))); }   /* layer.h:81-83 */

/* glibc rand() replay is not needed here: the test passes the parameters in.  Deterministic filler for standalone use:
 * the reference's law 0.5 - U[0,1) (layer.h:49-52) from a small LCG, per neuron bias first then its weights. */
void l5_init_params(float *p, uint32_t seed) {
    uint32_t s = seed ? seed : 1u;
    float tmp[L5_NPARAM];
    for (int i = 0; i < L5_NPARAM; ++i) {
        s = s * 1664525u + 1013904223u;
        tmp[i] = 0.5f - (float)(s >> 8) / 16777216.0f;
    }
    memcpy(p, tmp, sizeof(tmp));
}

void l5_forward(const float *p, const float *img, l5_acts *a) {
    const float (*in)[28] = (const float (*)[28])img;
    for (int m = 0; m < 6; ++m)
        for (int x = 0; x < 24; ++x)
            for (int y = 0; y < 24; ++y) {
                float s = 0.0f;
                for (int i = 0; i < 5; ++i)
                    for (int j = 0; j < 5; ++j) s += p[L5_C1W + m * 25 + i * 5 + j] * in[x + i][y + j];
                a->c1_pre[m][x][y] = s + p[L5_C1B + m];
                a->c1_out[m][x][y] = sigm(a->c1_pre[m][x][y]);
            }
    for (int m = 0; m < 6; ++m)
        for (int x = 0; x < 12; ++x)
            for (int y = 0; y < 12; ++y) {
                float s = 0.0f;
                for (int i = 0; i < 2; ++i)
                    for (int j = 0; j < 2; ++j) s += p[L5_S2W + i * 2 + j] * a->c1_out[m][2 * x + i][2 * y + j];
                a->s2_pre[m][x][y] = s + p[L5_S2B];
                a->s2_out[m][x][y] = sigm(a->s2_pre[m][x][y]);
            }
    for (int k = 0; k < 16; ++k)
        for (int x = 0; x < 8; ++x)
            for (int y = 0; y < 8; ++y) {
                float s = 0.0f;
                for (int c = 0; c < 6; ++c)
                    for (int i = 0; i < 5; ++i)
                        for (int j = 0; j < 5; ++j) s += p[L5_C3W + ((k * 6 + c) * 5 + i) * 5 + j] * a->s2_out[c][x + i][y + j];
                a->c3_pre[k][x][y] = s + p[L5_C3B + k];
                a->c3_out[k][x][y] = sigm(a->c3_pre[k][x][y]);
            }
    for (int m = 0; m < 16; ++m)
        for (int x = 0; x < 4; ++x)
            for (int y = 0; y < 4; ++y) {
                float s = 0.0f;
                for (int i = 0; i < 2; ++i)
                    for (int j = 0; j < 2; ++j) s += p[L5_S4W + i * 2 + j] * a->c3_out[m][2 * x + i][2 * y + j];
                a->s4_pre[m][x][y] = s + p[L5_S4B];
                a->s4_out[m][x][y] = sigm(a->s4_pre[m][x][y]);
            }
    const float *s4 = &a->s4_out[0][0][0];
    for (int o = 0; o < 10; ++o) {
        float s = 0.0f;
        for (int k = 0; k < 256; ++k) s += p[L5_FW + o * 256 + k] * s4[k];
        a->f_pre[o] = s + p[L5_FB + o];
        a->f_out[o] = sigm(a->f_pre[o]);
    }
}

/* packed (negative) gradient of one sample into g[L5_NPARAM] (assigned, not accumulated); returns the error norm
 * sqrt(sum d^2) of Main.cpp:28-34 */
float l5_backward(const float *p, const float *img, unsigned label, const l5_acts *a, float *g) {
    const float (*in)[28] = (const float (*)[28])img;
    float d_f[10], dpre_s4[16][4][4], dpre_c3[16][8][8], dout_s2[6][12][12], dpre_s2[6][12][12], dpre_c1[6][24][24];
    float ss = 0.0f;
    for (int o = 0; o < 10; ++o) {
        d_f[o] = (o == (int)label ? 1.0f : 0.0f) - a->f_out[o];
        ss += d_f[o] * d_f[o];
        g[L5_FB + o] = d_f[o];
    }
    const float *s4 = &a->s4_out[0][0][0];
    for (int o = 0; o < 10; ++o)
        for (int k = 0; k < 256; ++k) g[L5_FW + o * 256 + k] = d_f[o] * s4[k];
    float bs4 = 0.0f;
    for (int k = 0; k < 256; ++k) {
        float d = 0.0f;
        for (int o = 0; o < 10; ++o) d += p[L5_FW + o * 256 + k] * d_f[o];
        const float ov = s4[k];
        (&dpre_s4[0][0][0])[k] = d * ov * (1.0f - ov);
        bs4 += (&dpre_s4[0][0][0])[k];
    }
    g[L5_S4B] = bs4;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            float s = 0.0f;
            for (int m = 0; m < 16; ++m)
                for (int x = 0; x < 4; ++x)
                    for (int y = 0; y < 4; ++y) s += dpre_s4[m][x][y] * a->c3_out[m][2 * x + i][2 * y + j];
            g[L5_S4W + i * 2 + j] = s;
        }
    for (int m = 0; m < 16; ++m) {
        float bs = 0.0f;
        for (int x = 0; x < 8; ++x)
            for (int y = 0; y < 8; ++y) {
                const float dout = p[L5_S4W + (x & 1) * 2 + (y & 1)] * dpre_s4[m][x >> 1][y >> 1];
                const float o = a->c3_out[m][x][y];
                dpre_c3[m][x][y] = dout * (o * (1.0f - o));
                bs += dpre_c3[m][x][y];
            }
        g[L5_C3B + m] = bs;
    }
    for (int k = 0; k < 16; ++k)
        for (int c = 0; c < 6; ++c)
            for (int i = 0; i < 5; ++i)
                for (int j = 0; j < 5; ++j) {
                    float s = 0.0f;
                    for (int x = 0; x < 8; ++x)
                        for (int y = 0; y < 8; ++y) s += dpre_c3[k][x][y] * a->s2_out[c][x + i][y + j] / 64.0f;
                    g[L5_C3W + ((k * 6 + c) * 5 + i) * 5 + j] = s;
                }
    memset(dout_s2, 0, sizeof(dout_s2));
    for (int k = 0; k < 16; ++k)
        for (int c = 0; c < 6; ++c)
            for (int x = 0; x < 8; ++x)
                for (int y = 0; y < 8; ++y)
                    for (int i = 0; i < 5; ++i)
                        for (int j = 0; j < 5; ++j) dout_s2[c][x + i][y + j] += p[L5_C3W + ((k * 6 + c) * 5 + i) * 5 + j] * dpre_c3[k][x][y];
    float bs2 = 0.0f;
    for (int c = 0; c < 6; ++c)
        for (int x = 0; x < 12; ++x)
            for (int y = 0; y < 12; ++y) {
                const float o = a->s2_out[c][x][y];
                dpre_s2[c][x][y] = dout_s2[c][x][y] * o * (1.0f - o);
                bs2 += dpre_s2[c][x][y];
            }
    g[L5_S2B] = bs2;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            float s = 0.0f;
            for (int m = 0; m < 6; ++m)
                for (int x = 0; x < 12; ++x)
                    for (int y = 0; y < 12; ++y) s += dpre_s2[m][x][y] * a->c1_out[m][2 * x + i][2 * y + j];
            g[L5_S2W + i * 2 + j] = s;
        }
    for (int m = 0; m < 6; ++m) {
        float bs = 0.0f;
        for (int x = 0; x < 24; ++x)
            for (int y = 0; y < 24; ++y) {
                const float dout = p[L5_S2W + (x & 1) * 2 + (y & 1)] * dpre_s2[m][x >> 1][y >> 1];
                const float o = a->c1_out[m][x][y];
                dpre_c1[m][x][y] = dout * (o * (1.0f - o));
                bs += dpre_c1[m][x][y];
            }
        g[L5_C1B + m] = bs;
    }
    for (int m = 0; m < 6; ++m)
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j) {
                float s = 0.0f;
                for (int x = 0; x < 24; ++x)
                    for (int y = 0; y < 24; ++y) s += dpre_c1[m][x][y] * in[x + i][y + j] / 576.0f;
                g[L5_C1W + m * 25 + i * 5 + j] = s;
            }
    return (float)sqrt((double)ss);
}

/* frozen-parameter batch: g (double) = sum over samples of the packed gradient, err_sum = sum of error norms */
void l5_batch_grad(const float *p, const float *imgs, const uint8_t *labels, long B, double *g, double *err_sum) {
    static l5_acts a;
    static float gs[L5_NPARAM];
    for (int i = 0; i < L5_NPARAM; ++i) g[i] = 0.0;
    *err_sum = 0.0;
    for (long b = 0; b < B; ++b) {
        l5_forward(p, imgs + b * 784, &a);
        *err_sum += (double)l5_backward(p, imgs + b * 784, labels[b], &a, gs);
        for (int i = 0; i < L5_NPARAM; ++i) g[i] += (double)gs[i];
    }
}

/* w += step * g with the bias divisors of the rules above (layer.h:99, :316, :412 operand order: step * g / n) */
void l5_apply_update(float *p, const float *g, float step) {
    for (int i = 0; i < L5_NPARAM; ++i) {
        float d = step * g[i];
        if (i >= L5_C1B && i < L5_S2W) d = d / 576.0f;
        else if (i == L5_S2B) d = d / 864.0f;
        else if (i >= L5_C3B && i < L5_S4W) d = d / 64.0f;
        else if (i == L5_S4B) d = d / 256.0f;
        p[i] += d;
    }
}

void l5_forward_out(const float *p, const float *img, float *f_out10) {
    static l5_acts a;
    l5_forward(p, img, &a);
    memcpy(f_out10, a.f_out, sizeof(a.f_out));
}
