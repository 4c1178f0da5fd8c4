/*
 * Plain-C caller of libpamr_b200.so through include/pamr_b200.h: no PyTorch, no C++.
 * Host buffers in, uint8 label map out (pamr_pseudo_labels_host_f32 = run_pamr -> _rescale_and_clean ->
 * pseudo_gtmask -> argmax of the reference, models/SoftMaxAE.py:176-179, 263-268, 29-50, 61-67).
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -L1-stage-wseg_b200 -lpamr_b200 -Wl,-rpath,$PWD/1-stage-wseg_b200 -o c_abi_demo
 *   ./c_abi_demo            (needs a B200; prints the label histogram of a synthetic batch)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "pamr_b200.h"

int main(void) {
    const int B = 2, K = 3, C = 21, H = 96, W = 112, h = 24, w = 28, iters = 10;
    const int dil[6] = {1, 2, 4, 8, 12, 24};
    float* img = (float*)malloc(sizeof(float) * B * K * H * W);
    float* msk = (float*)malloc(sizeof(float) * B * C * h * w);
    float* lab = (float*)malloc(sizeof(float) * B * (C - 1));
    uint8_t* out = (uint8_t*)malloc((size_t)B * H * W);
    unsigned s = 12345u;
    for (int i = 0; i < B * K * H * W; ++i) { s = s * 1664525u + 1013904223u; img[i] = (float)(s >> 8) / 16777216.0f; }
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < h * w; ++p) {
            float sum = 0.f;
            for (int c = 0; c < C; ++c) { s = s * 1664525u + 1013904223u; msk[(b * C + c) * h * w + p] = (float)(s >> 8) / 16777216.0f; sum += msk[(b * C + c) * h * w + p]; }
            for (int c = 0; c < C; ++c) msk[(b * C + c) * h * w + p] /= sum;
        }
    for (int i = 0; i < B * (C - 1); ++i) lab[i] = (i % 3 == 0) ? 1.f : 0.f;
    printf("libpamr_b200 ABI %d\n", pamr_b200_abi_version());
    const int rc = pamr_pseudo_labels_host_f32(img, msk, lab, out, B, K, C, H, W, h, w, dil, 6, iters, 0.7f, 0.6f, 0.2f, 0);
    if (rc != 0) {
        fprintf(stderr, "pamr_pseudo_labels_host_f32 failed (%d): %s\n", rc, pamr_last_error());
        return 1;
    }
    long hist[256] = {0};
    for (long i = 0; i < (long)B * H * W; ++i) hist[out[i]]++;
    for (int c = 0; c < 256; ++c)
        if (hist[c]) printf("label %3d: %ld pixels\n", c, hist[c]);
    printf("kernels launched: %llu\n", (unsigned long long)pamr_launch_count());
    free(img); free(msk); free(lab); free(out);
    return 0;
}
