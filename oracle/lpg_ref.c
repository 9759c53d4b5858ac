/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's TensorFlow custom op
 * `LocalPlanarGuidance` / `LocalPlanarGuidanceGrad` CPU functors
 * (tensorflow/custom_layer/local_planar_guidance.cc:74-115 forward, :241-298 gradient).
 *
 * Layout is the TF one: input (B,h,w,4) NHWC, depth (B,H,W).  `focal` is read by the reference
 * into an unused local (Q1) and is therefore not a parameter here.  The gradient reproduces the
 * reference's formula INCLUDING its omission of the factor n4 on d/dn1..n3 (Q5) when tf_compat=1;
 * tf_compat=0 gives the true gradient (what PyTorch autograd computes for pytorch/bts.py:132-146).
 *
 * The real op cannot be built here (TensorFlow absent); the functor bodies carry no
 * TF types, so this file pins the NHWC addressing and the Q5 arithmetic.  Only tests/ and bench.py's
 * cpu_baseline leg link it (as oracle/_ref-free `liblpg_oracle.so`). */
#include <stddef.h>

void lpg_oracle_fwd(const float *plane, float *depth, int B, int h, int w, int r)
{
    const int H = h * r, W = w * r;
    for (int b = 0; b < B; ++b)
        for (int row = 0; row < H; ++row)
            for (int col = 0; col < W; ++col) {
                const float v = ((float)(row % r) - (float)(r - 1.0f) / 2.0f) / (float)r;
                const float u = ((float)(col % r) - (float)(r - 1.0f) / 2.0f) / (float)r;
                const float *p = plane + (((size_t)b * h + row / r) * w + col / r) * 4;
                const float den = p[0] * u + p[1] * v + p[2];
                depth[((size_t)b * H + row) * W + col] = p[3] / den;
            }
}

void lpg_oracle_bwd(const float *dy, const float *plane, float *dplane,
                    int B, int h, int w, int r, int tf_compat)
{
    const int H = h * r, W = w * r;
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < h; ++i)
            for (int j = 0; j < w; ++j) {
                const size_t idx = ((size_t)b * h + i) * w + j;
                const float n1 = plane[idx * 4 + 0], n2 = plane[idx * 4 + 1];
                const float n3 = plane[idx * 4 + 2], n4 = plane[idx * 4 + 3];
                float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
                for (int a = 0; a < r; ++a)
                    for (int c = 0; c < r; ++c) {
                        const float v = ((float)a - (float)(r - 1.0f) / 2.0f) / (float)r;
                        const float u = ((float)c - (float)(r - 1.0f) / 2.0f) / (float)r;
                        const float g = dy[((size_t)b * H + i * r + a) * W + j * r + c];
                        const float den = n1 * u + n2 * v + n3;
                        const float den2 = den * den;
                        const float s = tf_compat ? 1.0f : n4;
                        g0 += g * s * (-1.0f * u) / den2;
                        g1 += g * s * (-1.0f * v) / den2;
                        g2 += g * s * (-1.0f) / den2;
                        g3 += g / den;
                    }
                dplane[idx * 4 + 0] = g0; dplane[idx * 4 + 1] = g1;
                dplane[idx * 4 + 2] = g2; dplane[idx * 4 + 3] = g3;
            }
}
