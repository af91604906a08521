/*
 * Oracle port — AlphaZero network forward in plain fp32 loops. TEST INFRASTRUCTURE ONLY (see mzo.h).
 *
 *   network/py/alphazero_network.py:90-113  AlphaZeroNetwork.forward       -> mzo_net_forward
 *   network/py/network_unit.py:6-23         ResidualBlock (conv-BN-ReLU-conv-BN-add-ReLU)
 *   network/py/network_unit.py:26-42        PolicyNetwork (conv1x1, BN, ReLU, fc)
 *   network/py/network_unit.py:45-65        ValueNetwork  (conv1x1, BN, ReLU, fc1, ReLU, fc2, tanh)
 *
 * BatchNorm is applied un-folded, in eval mode with running statistics and eps = 1e-5
 * (torch.nn.BatchNorm2d default), so that the CUDA path's BN folding is checked against the
 * un-folded definition. Parameters are addressed by their state_dict names. Slow (naive direct
 * convolution): meant for small nets / few positions; full-size nets are checked against the
 * TorchScript module itself.
 */
#include "mzo.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_PARAMS 512

typedef struct {
    char name[96];
    float* data;
    int64_t numel;
} param;

struct mzo_net {
    int c, h, w, hidden, blocks, actions, value_hidden, pol_ch;
    param params[MAX_PARAMS];
    int num_params;
};

mzo_net* mzo_net_create(int c, int h, int w, int hidden, int blocks, int actions, int value_hidden)
{
    mzo_net* n = (mzo_net*)calloc(1, sizeof(mzo_net));
    n->c = c, n->h = h, n->w = w, n->hidden = hidden, n->blocks = blocks, n->actions = actions, n->value_hidden = value_hidden;
    n->pol_ch = (actions + h * w - 1) / (h * w); /* network_unit.py:31 */
    return n;
}

void mzo_net_destroy(mzo_net* n)
{
    if (!n) { return; }
    for (int i = 0; i < n->num_params; ++i) { free(n->params[i].data); }
    free(n);
}

int mzo_net_set(mzo_net* n, const char* name, const float* data, int64_t numel)
{
    if (n->num_params >= MAX_PARAMS || strlen(name) >= sizeof(n->params[0].name)) { return -1; }
    param* p = &n->params[n->num_params++];
    strcpy(p->name, name);
    p->numel = numel;
    p->data = (float*)malloc(sizeof(float) * (size_t)numel);
    memcpy(p->data, data, sizeof(float) * (size_t)numel);
    return 0;
}

static const float* get(const mzo_net* n, const char* prefix, const char* suffix, int64_t expect)
{
    char key[160];
    snprintf(key, sizeof(key), "%s%s", prefix, suffix);
    for (int i = 0; i < n->num_params; ++i) {
        if (strcmp(n->params[i].name, key) == 0) {
            if (n->params[i].numel != expect) {
                fprintf(stderr, "mzo_net: %s has %lld elements, expected %lld\n", key, (long long)n->params[i].numel, (long long)expect);
                abort();
            }
            return n->params[i].data;
        }
    }
    fprintf(stderr, "mzo_net: missing parameter %s\n", key);
    abort();
}

/* out[co][y][x] = bias[co] + sum_{ci,ky,kx} w[co][ci][ky][kx] * in[ci][y+ky-p][x+kx-p]; then BN (eval) */
static void conv_bn(const mzo_net* n, const char* conv, const char* bn, int cin, int cout, int k, const float* in, float* out)
{
    int H = n->h, W = n->w, p = k / 2;
    const float* w = get(n, conv, ".weight", (int64_t)cout * cin * k * k);
    const float* b = get(n, conv, ".bias", cout);
    const float* g = get(n, bn, ".weight", cout);
    const float* be = get(n, bn, ".bias", cout);
    const float* mu = get(n, bn, ".running_mean", cout);
    const float* var = get(n, bn, ".running_var", cout);
    for (int co = 0; co < cout; ++co) {
        for (int y = 0; y < H; ++y) {
            for (int x = 0; x < W; ++x) {
                float acc = 0.0f;
                for (int ci = 0; ci < cin; ++ci) {
                    for (int ky = 0; ky < k; ++ky) {
                        int yy = y + ky - p;
                        if (yy < 0 || yy >= H) { continue; }
                        for (int kx = 0; kx < k; ++kx) {
                            int xx = x + kx - p;
                            if (xx < 0 || xx >= W) { continue; }
                            acc += w[((co * cin + ci) * k + ky) * k + kx] * in[(ci * H + yy) * W + xx];
                        }
                    }
                }
                acc += b[co];
                out[(co * H + y) * W + x] = (acc - mu[co]) / sqrtf(var[co] + 1e-5f) * g[co] + be[co];
            }
        }
    }
}

static void relu(float* x, int n)
{
    for (int i = 0; i < n; ++i) { x[i] = (x[i] > 0.0f ? x[i] : 0.0f); }
}

static void linear(const mzo_net* n, const char* fc, int in_f, int out_f, const float* in, float* out)
{
    const float* w = get(n, fc, ".weight", (int64_t)out_f * in_f);
    const float* b = get(n, fc, ".bias", out_f);
    for (int o = 0; o < out_f; ++o) {
        float acc = 0.0f;
        for (int i = 0; i < in_f; ++i) { acc += w[o * in_f + i] * in[i]; }
        out[o] = acc + b[o];
    }
}

void mzo_net_forward(const mzo_net* n, const float* features, int batch, float* policy, float* logits, float* value)
{
    int HW = n->h * n->w, Ch = n->hidden;
    float* x = (float*)malloc(sizeof(float) * (size_t)(Ch * HW));
    float* y = (float*)malloc(sizeof(float) * (size_t)(Ch * HW));
    float* z = (float*)malloc(sizeof(float) * (size_t)(Ch * HW));
    float* ph = (float*)malloc(sizeof(float) * (size_t)(n->pol_ch * HW));
    float* vh = (float*)malloc(sizeof(float) * (size_t)(HW + n->value_hidden));
    char c1[64], b1[64], c2[64], b2[64];
    for (int s = 0; s < batch; ++s) {
        const float* in = features + (size_t)s * n->c * HW;
        conv_bn(n, "conv", "bn", n->c, Ch, 3, in, x); /* alphazero_network.py:91-93 */
        relu(x, Ch * HW);
        for (int blk = 0; blk < n->blocks; ++blk) { /* network_unit.py:14-23 */
            snprintf(c1, sizeof(c1), "residual_blocks.%d.conv1", blk);
            snprintf(b1, sizeof(b1), "residual_blocks.%d.bn1", blk);
            snprintf(c2, sizeof(c2), "residual_blocks.%d.conv2", blk);
            snprintf(b2, sizeof(b2), "residual_blocks.%d.bn2", blk);
            conv_bn(n, c1, b1, Ch, Ch, 3, x, y);
            relu(y, Ch * HW);
            conv_bn(n, c2, b2, Ch, Ch, 3, y, z);
            for (int i = 0; i < Ch * HW; ++i) { x[i] = x[i] + z[i]; }
            relu(x, Ch * HW);
        }
        /* policy head, network_unit.py:36-42; softmax alphazero_network.py:99 */
        conv_bn(n, "policy.conv", "policy.bn", Ch, n->pol_ch, 1, x, ph);
        relu(ph, n->pol_ch * HW);
        float* lg = logits + (size_t)s * n->actions;
        linear(n, "policy.fc", n->pol_ch * HW, n->actions, ph, lg);
        float mx = lg[0], sum = 0.0f;
        for (int a = 1; a < n->actions; ++a) { mx = (lg[a] > mx ? lg[a] : mx); }
        for (int a = 0; a < n->actions; ++a) {
            policy[(size_t)s * n->actions + a] = expf(lg[a] - mx);
            sum += policy[(size_t)s * n->actions + a];
        }
        for (int a = 0; a < n->actions; ++a) { policy[(size_t)s * n->actions + a] /= sum; }
        /* value head, network_unit.py:56-65 */
        conv_bn(n, "value.conv", "value.bn", Ch, 1, 1, x, vh);
        relu(vh, HW);
        linear(n, "value.fc1", HW, n->value_hidden, vh, vh + HW);
        relu(vh + HW, n->value_hidden);
        float v;
        linear(n, "value.fc2", n->value_hidden, 1, vh + HW, &v);
        value[s] = tanhf(v);
    }
    free(x), free(y), free(z), free(ph), free(vh);
}
