/*
 * oracle/hash_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's multiresolution hash-grid operator
 * (reference: hashencoder/src/hashencoder.cu).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may call into this file.  The product path
 * (holoscene_b200/csrc) never links or loads it.
 *
 * Parity pin: tests/golden/hash_ref_*.npz are outputs of the reference's own CUDA kernels
 * (oracle/_ref, built by oracle/build_ref.py) run on a B200; tests/test_oracle_hash.py checks this
 * file against them.
 *
 * Layouts follow the reference FFI (hashencoder/src/hashencoder.h:13-15):
 *   inputs  [B,3]  in [0,1]          embeddings [sum(n_l), C]        offsets [L+1] int32
 *   outputs [L,B,C]                  dy_dx [B, L*3*C]  (b, l, d, c)   grad [L,B,C]
 * D is fixed to 3 (the only value the Stage-1 path uses); C <= 8.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define HSO_D 3
#define HSO_MAXC 8

/* reference: hashencoder.cu:36-51 (xor of coordinate * prime, uint32 wrap-around) */
static inline uint32_t hso_hash(const uint32_t g[HSO_D]) {
    return (g[0] * 1u) ^ (g[1] * 2654435761u) ^ (g[2] * 805459861u);
}

/* reference: hashencoder.cu:54-72 -- dense index while the running stride still fits, else hash.
 * Returns the ROW (not multiplied by C). */
static inline uint32_t hso_row(uint32_t hashmap_size, uint32_t resolution, const uint32_t g[HSO_D]) {
    uint32_t stride = 1, index = 0;
    for (int d = 0; d < HSO_D && stride <= hashmap_size; ++d) {
        index += g[d] * stride;
        stride *= resolution;
    }
    if (stride > hashmap_size) index = hso_hash(g);
    return index % hashmap_size;
}

typedef struct {
    int oob;
    uint32_t hashmap_size, resolution;
    float scale;
    float w[HSO_D];   /* smoothstep(frac)            hashencoder.cu:87-89  */
    float dw[HSO_D];  /* smoothstep'(frac) = 6t(1-t) hashencoder.cu:91-93  */
    uint32_t g[HSO_D];
} hso_cell;

/* reference: hashencoder.cu:124-167 (range test, per-level scale/resolution, cell + fraction) */
static inline void hso_locate(const float* x, const int* offsets, uint32_t level, float S, uint32_t H, hso_cell* c) {
    c->oob = 0;
    for (int d = 0; d < HSO_D; ++d)
        if (x[d] < 0.0f || x[d] > 1.0f) c->oob = 1;
    c->hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    c->scale = exp2f((float)level * S) * (float)H - 1.0f;
    c->resolution = (uint32_t)ceilf(c->scale) + 1u;
    if (c->oob) return;
    for (int d = 0; d < HSO_D; ++d) {
        float pos = x[d] * c->scale;
        float fl = floorf(pos);
        c->g[d] = (uint32_t)fl;
        float t = pos - (float)c->g[d];
        c->dw[d] = 6.0f * t * (1.0f - t);
        c->w[d] = t * t * (3.0f - 2.0f * t);
    }
}

/* reference: kernel_grid, hashencoder.cu:103-254 */
void hso_forward(const float* inputs, const float* grid, const int* offsets, float* outputs,
                 uint32_t B, uint32_t C, uint32_t L, float S, uint32_t H,
                 int calc_grad_inputs, float* dy_dx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (uint32_t level = 0; level < L; ++level) {
        for (uint32_t b = 0; b < B; ++b) {
            const float* tab = grid + (size_t)(uint32_t)offsets[level] * C;
            float* out = outputs + ((size_t)level * B + b) * C;
            float* dd = calc_grad_inputs ? dy_dx + (size_t)b * HSO_D * L * C + (size_t)level * HSO_D * C : 0;
            hso_cell c;
            hso_locate(inputs + (size_t)b * HSO_D, offsets, level, S, H, &c);
            if (c.oob) {
                for (uint32_t ch = 0; ch < C; ++ch) out[ch] = 0.0f;
                if (dd) for (uint32_t i = 0; i < HSO_D * C; ++i) dd[i] = 0.0f;
                continue;
            }
            float res[HSO_MAXC] = {0};
            for (uint32_t idx = 0; idx < 8u; ++idx) {
                float w = 1.0f;
                uint32_t gl[HSO_D];
                for (int d = 0; d < HSO_D; ++d) {
                    if ((idx & (1u << d)) == 0) { w *= 1.0f - c.w[d]; gl[d] = c.g[d]; }
                    else                        { w *= c.w[d];        gl[d] = c.g[d] + 1u; }
                }
                const float* e = tab + (size_t)hso_row(c.hashmap_size, c.resolution, gl) * C;
                for (uint32_t ch = 0; ch < C; ++ch) res[ch] += w * e[ch];
            }
            for (uint32_t ch = 0; ch < C; ++ch) out[ch] = res[ch];
            if (!dd) continue;
            /* hashencoder.cu:210-253: d/dx_gd = scale * sum_{4 corners of other axes} w_other*(right-left)*smoothstep' */
            for (int gd = 0; gd < HSO_D; ++gd) {
                float rg[HSO_MAXC] = {0};
                for (uint32_t idx = 0; idx < 4u; ++idx) {
                    float w = c.scale;
                    uint32_t gl[HSO_D];
                    for (int nd = 0; nd < HSO_D - 1; ++nd) {
                        int d = (nd >= gd) ? nd + 1 : nd;
                        if ((idx & (1u << nd)) == 0) { w *= 1.0f - c.w[d]; gl[d] = c.g[d]; }
                        else                         { w *= c.w[d];        gl[d] = c.g[d] + 1u; }
                    }
                    gl[gd] = c.g[gd];
                    const float* el = tab + (size_t)hso_row(c.hashmap_size, c.resolution, gl) * C;
                    gl[gd] = c.g[gd] + 1u;
                    const float* er = tab + (size_t)hso_row(c.hashmap_size, c.resolution, gl) * C;
                    for (uint32_t ch = 0; ch < C; ++ch) rg[ch] += w * (er[ch] - el[ch]) * c.dw[gd];
                }
                for (uint32_t ch = 0; ch < C; ++ch) dd[gd * C + ch] = rg[ch];
            }
        }
    }
}

/* reference: kernel_grid_backward (hashencoder.cu:257-343) + kernel_input_backward (:346-372).
 * grad_grid is ACCUMULATED into (caller pre-zeroes, hashgrid.py:76). Each level owns a disjoint
 * slice of grad_grid, so levels run in parallel without atomics; within a level the points are
 * visited in index order (deterministic, unlike the reference's atomics). */
void hso_backward(const float* grad, const float* inputs, const float* grid, const int* offsets,
                  float* grad_grid, uint32_t B, uint32_t C, uint32_t L, float S, uint32_t H,
                  int calc_grad_inputs, const float* dy_dx, float* grad_inputs) {
    (void)grid;
#pragma omp parallel for schedule(dynamic, 1)
    for (uint32_t level = 0; level < L; ++level) {
        float* gtab = grad_grid + (size_t)(uint32_t)offsets[level] * C;
        for (uint32_t b = 0; b < B; ++b) {
            hso_cell c;
            hso_locate(inputs + (size_t)b * HSO_D, offsets, level, S, H, &c);
            if (c.oob) continue;
            const float* g = grad + ((size_t)level * B + b) * C;
            for (uint32_t idx = 0; idx < 8u; ++idx) {
                float w = 1.0f;
                uint32_t gl[HSO_D];
                for (int d = 0; d < HSO_D; ++d) {
                    if ((idx & (1u << d)) == 0) { w *= 1.0f - c.w[d]; gl[d] = c.g[d]; }
                    else                        { w *= c.w[d];        gl[d] = c.g[d] + 1u; }
                }
                float* e = gtab + (size_t)hso_row(c.hashmap_size, c.resolution, gl) * C;
                for (uint32_t ch = 0; ch < C; ++ch) e[ch] += w * g[ch];
            }
        }
    }
    if (!calc_grad_inputs) return;
#pragma omp parallel for schedule(static)
    for (uint32_t b = 0; b < B; ++b) {
        const float* dd = dy_dx + (size_t)b * L * HSO_D * C;
        for (int d = 0; d < HSO_D; ++d) {
            float r = 0.0f;
            for (uint32_t l = 0; l < L; ++l)
                for (uint32_t ch = 0; ch < C; ++ch)
                    r += grad[((size_t)l * B + b) * C + ch] * dd[l * HSO_D * C + d * C + ch];
            grad_inputs[(size_t)b * HSO_D + d] = r;
        }
    }
}

/* reference: kernel_grid_second_backward_grad (hashencoder.cu:375-428) and
 * kernel_grid_second_backward_embedding (:431-595).  No d/d(inputs) term exists in the reference
 * (hashgrid.py:101 returns None for inputs) and none is produced here. */
void hso_second_backward(const float* grad, const float* inputs, const float* grid, const int* offsets,
                         uint32_t B, uint32_t C, uint32_t L, float S, uint32_t H,
                         const float* dy_dx, const float* grad_grad_inputs,
                         float* grad_grad, float* grad2_grid) {
    (void)grid;
#pragma omp parallel for schedule(dynamic, 1)
    for (uint32_t level = 0; level < L; ++level) {
        float* gtab = grad2_grid + (size_t)(uint32_t)offsets[level] * C;
        for (uint32_t b = 0; b < B; ++b) {
            const float* ggx = grad_grad_inputs + (size_t)b * HSO_D;
            const float* g = grad + ((size_t)level * B + b) * C;
            const float* dd = dy_dx + (size_t)b * L * HSO_D * C + (size_t)level * HSO_D * C;
            float* gg = grad_grad + ((size_t)level * B + b) * C;
            for (uint32_t ch = 0; ch < C; ++ch) {
                float r = 0.0f;
                for (int d = 0; d < HSO_D; ++d) r += ggx[d] * dd[d * C + ch];
                gg[ch] = r;
            }
            hso_cell c;
            hso_locate(inputs + (size_t)b * HSO_D, offsets, level, S, H, &c);
            if (c.oob) continue;
            float cache[8][HSO_MAXC];
            memset(cache, 0, sizeof(cache));
            for (int gd = 0; gd < HSO_D; ++gd) {
                for (uint32_t idx = 0; idx < 4u; ++idx) {
                    float w = c.scale;
                    uint32_t bits = 0;
                    for (int nd = 0; nd < HSO_D - 1; ++nd) {
                        int d = (nd >= gd) ? nd + 1 : nd;
                        if ((idx & (1u << nd)) == 0) { w *= 1.0f - c.w[d]; }
                        else                         { w *= c.w[d]; bits |= (1u << d); }
                    }
                    uint32_t left = bits, right = bits | (1u << gd);
                    for (uint32_t ch = 0; ch < C; ++ch) {
                        float v = w * g[ch] * ggx[gd] * c.dw[gd];
                        cache[right][ch] += v;
                        cache[left][ch] -= v;
                    }
                }
            }
            for (uint32_t idx = 0; idx < 8u; ++idx) {
                uint32_t gl[HSO_D];
                for (int d = 0; d < HSO_D; ++d) gl[d] = c.g[d] + ((idx >> d) & 1u);
                float* e = gtab + (size_t)hso_row(c.hashmap_size, c.resolution, gl) * C;
                for (uint32_t ch = 0; ch < C; ++ch) e[ch] += cache[idx][ch];
            }
        }
    }
}
