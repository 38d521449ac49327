/*
 * CPU ORACLE (C restatement) -- TEST INFRASTRUCTURE ONLY.
 *
 * Restates /root/reference/src/attention_ref.zig in plain C with the SAME
 * loop nest and fp32 accumulation order, because the Zig original cannot be
 * compiled in this image (no zig toolchain).  Only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 *   aule_ref_forward         attention_ref.zig:18-93    (AttentionRef.forward)
 *   aule_ref_forward_causal  attention_ref.zig:97-171   (AttentionRef.forwardCausal)
 *   aule_ref_compare_arrays  attention_ref.zig:175-184  (compareArrays)
 *   aule_ref_max_abs_diff    attention_ref.zig:187-196  (maxAbsDiff)
 *   aule_ref_mean_abs_diff   attention_ref.zig:199-206  (meanAbsDiff)
 *
 * Pinned by the reference's own known-answer vectors (attention_ref.zig:250-298,
 * tests/test_attention.zig:158-270) in tests/test_oracle.py, and cross-checked
 * there against the NumPy restatement and the committed golden outputs of the
 * reference's NumPy path.
 *
 * Build: make -C oracle   (gcc -O2 -fno-fast-math: keep IEEE fp32 order)
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

static int ref_forward_impl(const float* Q, const float* K, const float* V, float* out,
                            size_t batch_size, size_t num_heads, size_t seq_len,
                            size_t head_dim, int causal) {
    /* attention_ref.zig:29 / :108 */
    const float scale = 1.0f / sqrtf((float)head_dim);

    /* scratch S x S scores, attention_ref.zig:32 / :110 */
    float* scores = (float*)malloc(sizeof(float) * seq_len * seq_len);
    if (!scores) return -2;

    for (size_t b = 0; b < batch_size; ++b) {
        for (size_t h = 0; h < num_heads; ++h) {
            const size_t base = (b * num_heads + h) * seq_len * head_dim;

            /* Step 1: S[i,j] = (sum_d Q[i,d] K[j,d]) * scale ; causal: j > i -> -inf
             * attention_ref.zig:42-53 and :116-131 */
            for (size_t i = 0; i < seq_len; ++i) {
                for (size_t j = 0; j < seq_len; ++j) {
                    if (causal && j > i) {
                        scores[i * seq_len + j] = -INFINITY;
                    } else {
                        float dot = 0.0f;
                        for (size_t d = 0; d < head_dim; ++d) {
                            dot += Q[base + i * head_dim + d] * K[base + j * head_dim + d];
                        }
                        scores[i * seq_len + j] = dot * scale;
                    }
                }
            }

            /* Step 2: row softmax, three passes (max, exp+sum, normalise)
             * attention_ref.zig:56-78 and :134-153 */
            for (size_t i = 0; i < seq_len; ++i) {
                float* row = scores + i * seq_len;
                float row_max = -INFINITY;
                for (size_t j = 0; j < seq_len; ++j) row_max = fmaxf(row_max, row[j]);
                float row_sum = 0.0f;
                for (size_t j = 0; j < seq_len; ++j) {
                    const float e = expf(row[j] - row_max);
                    row[j] = e;
                    row_sum += e;
                }
                const float inv_sum = 1.0f / row_sum;
                for (size_t j = 0; j < seq_len; ++j) row[j] *= inv_sum;
            }

            /* Step 3: O[i,d] = sum_j P[i,j] V[j,d]   attention_ref.zig:81-91 / :156-167 */
            for (size_t i = 0; i < seq_len; ++i) {
                for (size_t d = 0; d < head_dim; ++d) {
                    float acc = 0.0f;
                    for (size_t j = 0; j < seq_len; ++j) {
                        acc += scores[i * seq_len + j] * V[base + j * head_dim + d];
                    }
                    out[base + i * head_dim + d] = acc;
                }
            }
        }
    }
    free(scores);
    return 0;
}

int aule_ref_forward(const float* Q, const float* K, const float* V, float* out,
                     uint32_t B, uint32_t H, uint32_t S, uint32_t D) {
    return ref_forward_impl(Q, K, V, out, B, H, S, D, 0);
}

int aule_ref_forward_causal(const float* Q, const float* K, const float* V, float* out,
                            uint32_t B, uint32_t H, uint32_t S, uint32_t D) {
    return ref_forward_impl(Q, K, V, out, B, H, S, D, 1);
}

/* attention_ref.zig:175-184 */
int aule_ref_compare_arrays(const float* expected, const float* actual, size_t n, float tol) {
    for (size_t i = 0; i < n; ++i)
        if (fabsf(expected[i] - actual[i]) > tol) return 0;
    return 1;
}

/* attention_ref.zig:187-196 */
float aule_ref_max_abs_diff(const float* a, const float* b, size_t n) {
    float m = 0.0f;
    for (size_t i = 0; i < n; ++i) m = fmaxf(m, fabsf(a[i] - b[i]));
    return m;
}

/* attention_ref.zig:199-206 */
float aule_ref_mean_abs_diff(const float* a, const float* b, size_t n) {
    if (n == 0) return INFINITY;
    float s = 0.0f;
    for (size_t i = 0; i < n; ++i) s += fabsf(a[i] - b[i]);
    return s / (float)n;
}
