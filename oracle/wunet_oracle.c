/*
 * wunet_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference Wave-U-Net forward (eval-mode BatchNorm)
 * in plain C.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call into this file; the product
 * path (wave_u_net_for_speech_enhancement_b200/) never does.
 *
 * Follows /root/reference/model/unet_basic.py line by line:
 *   DownSamplingLayer  :6-17   Conv1d(k=15,p=7) -> BatchNorm1d -> LeakyReLU(0.1)
 *   UpSamplingLayer    :19-30  Conv1d(k=5,p=2)  -> BatchNorm1d -> LeakyReLU(0.1)
 *   Model.__init__     :33-75  channel plan
 *   Model.forward      :77-100 encoder / decimate / middle / upsample / concat /
 *                              decoder / concat input / 1x1 conv / tanh
 * The arithmetic that PyTorch supplies (not in the reference tree) is restated
 * from its published definition:
 *   Conv1d     y[b,co,l] = bias[co] + sum_ci sum_k W[co,ci,k] x[b,ci,l+k-pad], zero padded
 *   BatchNorm1d (eval) y = (x - running_mean) / sqrt(running_var + 1e-5) * weight + bias
 *   LeakyReLU  y = x >= 0 ? x : 0.1 x
 *   F.interpolate(scale_factor=2, mode="linear", align_corners=True):
 *              scale = (Lin-1)/(Lout-1) (fp32), src = scale*i (fp32), i0=(int)src,
 *              i1 = i0 + (i0 < Lin-1), lam = src - i0, y = (1-lam) x[i0] + lam x[i1]
 * Parity is pinned by tests/golden/ (generated from the live reference module by
 * oracle/gen_golden.py); see tests/test_oracle.py.
 *
 * Accumulation is in double, storage between layers in float (the reference
 * stores fp32 activations), so this oracle is at least as accurate as the
 * reference's own fp32 CPU path (measured difference ~1e-6).
 *
 * Parameter packing ("flat" layout, all float32), in state_dict order:
 *   for blk in encoder[0..n-1], middle, decoder[0..n-1]:
 *       W[Cout*Cin*K], bias[Cout], bn_weight[Cout], bn_bias[Cout],
 *       bn_running_mean[Cout], bn_running_var[Cout]
 *   out.W[1*(ci+1)*1], out.bias[1]
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <pthread.h>
#include <unistd.h>

#define WUNET_ORACLE_MAX_LAYERS 16

typedef struct {
    int cin, cout, k, pad;
    const float *w, *b, *g, *beta, *mean, *var;
} blk_t;

/* reference model/unet_basic.py:38-39 (encoder), :59-62 (decoder) */
static int plan(int n, int ci, const float *params, blk_t *blks /* 2n+1 */, const float **out_w,
                const float **out_b)
{
    const float *p = params;
    int nb = 0;
    for (int i = 0; i < n; ++i) {
        blk_t *q = &blks[nb++];
        q->cin = (i == 0) ? 1 : i * ci;
        q->cout = (i + 1) * ci;
        q->k = 15;
        q->pad = 7;
    }
    {
        blk_t *q = &blks[nb++];
        q->cin = n * ci;
        q->cout = n * ci;
        q->k = 15;
        q->pad = 7;
    }
    for (int j = 0; j < n; ++j) {
        blk_t *q = &blks[nb++];
        /* decoder_in = reversed([(2i+1)*ci for i in 1..n-1] + [2*n*ci]) */
        q->cin = (j == 0) ? 2 * n * ci : (2 * (n - j) + 1) * ci;
        q->cout = (n - j) * ci;
        q->k = 5;
        q->pad = 2;
    }
    for (int i = 0; i < nb; ++i) {
        blk_t *q = &blks[i];
        size_t wn = (size_t)q->cout * q->cin * q->k;
        q->w = p; p += wn;
        q->b = p; p += q->cout;
        q->g = p; p += q->cout;
        q->beta = p; p += q->cout;
        q->mean = p; p += q->cout;
        q->var = p; p += q->cout;
    }
    *out_w = p; p += (ci + 1);
    *out_b = p; p += 1;
    return nb;
}

size_t wunet_oracle_param_count(int n, int ci)
{
    blk_t blks[2 * WUNET_ORACLE_MAX_LAYERS + 1];
    const float *ow, *ob;
    float dummy = 0.f;
    if (n < 1 || n > WUNET_ORACLE_MAX_LAYERS) return 0;
    plan(n, ci, &dummy, blks, &ow, &ob);
    return (size_t)((ob + 1) - &dummy);
}

/* Conv1d + eval BatchNorm + LeakyReLU(0.1); x [B,cin,L], y [B,cout,L] (NCL, like the reference).
 * Work items are (b, cout) rows, handed out to pthreads through an atomic counter. */
typedef struct {
    const blk_t *q;
    int B, L;
    const float *x;
    float *y;
    int next; /* atomic work counter */
} conv_job_t;

static void *conv_worker(void *arg)
{
    conv_job_t *job = (conv_job_t *)arg;
    const blk_t *q = job->q;
    const int cin = q->cin, cout = q->cout, K = q->k, pad = q->pad, L = job->L;
    double *acc = (double *)malloc(sizeof(double) * (size_t)L);
    const int total = job->B * cout;
    for (;;) {
        const int item = __atomic_fetch_add(&job->next, 1, __ATOMIC_RELAXED);
        if (item >= total) break;
        const int b = item / cout, co = item % cout;
        const double bias = q->b[co];
        for (int l = 0; l < L; ++l) acc[l] = bias;
        for (int c = 0; c < cin; ++c) {
            const float *xr = job->x + ((size_t)b * cin + c) * L;
            const float *wr = q->w + ((size_t)co * cin + c) * K;
            for (int k = 0; k < K; ++k) {
                const double w = wr[k];
                const int off = k - pad;
                const int lo = off < 0 ? -off : 0;
                const int hi = off > 0 ? L - off : L;
                for (int l = lo; l < hi; ++l) acc[l] += w * (double)xr[l + off];
            }
        }
        /* BatchNorm1d eval: eps = 1e-5 (torch default), then LeakyReLU(0.1) */
        const double s = (double)q->g[co] / sqrt((double)q->var[co] + 1e-5);
        const double t = (double)q->beta[co] - (double)q->mean[co] * s;
        float *yr = job->y + ((size_t)b * cout + co) * L;
        for (int l = 0; l < L; ++l) {
            const double v = acc[l] * s + t;
            yr[l] = (float)(v >= 0.0 ? v : 0.1 * v);
        }
    }
    free(acc);
    return NULL;
}

static int g_threads = 0;
void wunet_oracle_set_threads(int n) { g_threads = n; }

static void conv_bn_lrelu(const blk_t *q, int B, int L, const float *x, float *y)
{
    conv_job_t job = {q, B, L, x, y, 0};
    int nt = g_threads > 0 ? g_threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nt > 256) nt = 256;
    if (nt < 1) nt = 1;
    if (nt > B * q->cout) nt = B * q->cout;
    pthread_t th[256];
    for (int i = 1; i < nt; ++i) pthread_create(&th[i], NULL, conv_worker, &job);
    conv_worker(&job);
    for (int i = 1; i < nt; ++i) pthread_join(th[i], NULL);
}

/*
 * levels_out: NULL, or an array of 2n+1 pointers (entries may be NULL); entry i
 * receives the full-resolution output of block i (encoder i / middle / decoder j)
 * in NCL fp32, for per-level parity checks.
 * returns 0 on success, nonzero on bad arguments.
 */
int wunet_oracle_forward(int n, int ci, int B, int T, const float *params, const float *x, float *y,
                         float *const *levels_out)
{
    if (n < 1 || n > WUNET_ORACLE_MAX_LAYERS || ci < 1 || B < 1 || T < 1) return 1;
    if (T % (1 << n) != 0) return 2; /* reference: torch.cat size error at unet_basic.py:95 */
    blk_t blks[2 * WUNET_ORACLE_MAX_LAYERS + 1];
    const float *ow, *ob;
    plan(n, ci, params, blks, &ow, &ob);

    float *skip[WUNET_ORACLE_MAX_LAYERS];
    /* encoder: unet_basic.py:82-86 */
    const float *cur = x;
    float *owned = NULL;
    int L = T;
    for (int i = 0; i < n; ++i) {
        const blk_t *q = &blks[i];
        skip[i] = (float *)malloc(sizeof(float) * (size_t)B * q->cout * L);
        conv_bn_lrelu(q, B, L, cur, skip[i]);
        if (levels_out && levels_out[i]) memcpy(levels_out[i], skip[i], sizeof(float) * (size_t)B * q->cout * L);
        /* o = o[:, :, ::2] */
        float *dec = (float *)malloc(sizeof(float) * (size_t)B * q->cout * (L / 2));
        for (size_t r = 0; r < (size_t)B * q->cout; ++r)
            for (int l = 0; l < L / 2; ++l) dec[r * (L / 2) + l] = skip[i][r * L + 2 * l];
        free(owned);
        owned = dec;
        cur = dec;
        L /= 2;
    }
    /* middle: unet_basic.py:88 */
    {
        const blk_t *q = &blks[n];
        float *m = (float *)malloc(sizeof(float) * (size_t)B * q->cout * L);
        conv_bn_lrelu(q, B, L, cur, m);
        if (levels_out && levels_out[n]) memcpy(levels_out[n], m, sizeof(float) * (size_t)B * q->cout * L);
        free(owned);
        owned = m;
        cur = m;
    }
    /* decoder: unet_basic.py:91-96 */
    int cprev = blks[n].cout;
    for (int j = 0; j < n; ++j) {
        const blk_t *q = &blks[n + 1 + j];
        const int e = n - 1 - j;
        const int cskip = blks[e].cout;
        const int Lin = L, Lout = 2 * L;
        float *cat = (float *)malloc(sizeof(float) * (size_t)B * (cprev + cskip) * Lout);
        /* F.interpolate(..., scale_factor=2, mode="linear", align_corners=True) in fp32 like ATen */
        const float scale = (Lout > 1) ? (float)(Lin - 1) / (float)(Lout - 1) : 0.f;
        for (int b = 0; b < B; ++b) {
            for (int c = 0; c < cprev; ++c) {
                const float *src = cur + ((size_t)b * cprev + c) * Lin;
                float *dst = cat + ((size_t)b * (cprev + cskip) + c) * Lout;
                for (int i = 0; i < Lout; ++i) {
                    const float s = scale * (float)i;
                    const int i0 = (int)s;
                    const int i1 = i0 + (i0 < Lin - 1 ? 1 : 0);
                    const float lam1 = s - (float)i0;
                    const float lam0 = 1.f - lam1;
                    dst[i] = lam0 * src[i0] + lam1 * src[i1];
                }
            }
            /* torch.cat([o, tmp[n-1-j]], dim=1) — upsampled first, then skip */
            memcpy(cat + ((size_t)b * (cprev + cskip) + cprev) * Lout, skip[e] + (size_t)b * cskip * Lout,
                   sizeof(float) * (size_t)cskip * Lout);
        }
        float *o = (float *)malloc(sizeof(float) * (size_t)B * q->cout * Lout);
        conv_bn_lrelu(q, B, Lout, cat, o);
        if (levels_out && levels_out[n + 1 + j])
            memcpy(levels_out[n + 1 + j], o, sizeof(float) * (size_t)B * q->cout * Lout);
        free(cat);
        free(owned);
        owned = o;
        cur = o;
        cprev = q->cout;
        L = Lout;
    }
    /* out: cat([o, input]) -> Conv1d(ci+1 -> 1, k=1) -> Tanh   (unet_basic.py:98-99) */
    for (int b = 0; b < B; ++b) {
        for (int l = 0; l < T; ++l) {
            double a = ob[0];
            for (int c = 0; c < cprev; ++c) a += (double)ow[c] * (double)cur[((size_t)b * cprev + c) * T + l];
            a += (double)ow[cprev] * (double)x[(size_t)b * T + l];
            y[(size_t)b * T + l] = (float)tanh(a);
        }
    }
    free(owned);
    for (int i = 0; i < n; ++i) free(skip[i]);
    return 0;
}
