/*
 * roundtrip.c -- the C ABI (include/lumacu.h) from plain C99: build the PQ-11 table the way
 * LumaQuantizer::setQuantizer does, encode one synthetic frame to 4:2:0 LE16 planes, decode it again and report
 * the worst relative round-trip error (quantisation only).
 *
 *   gcc -std=c99 -O2 -Iinclude examples/roundtrip.c -Llumahdrv_b200 -llumacu -Wl,-rpath,lumahdrv_b200 -lm -o roundtrip
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lumacu.h"

#define CHECK(call)                                                                                   \
    do {                                                                                              \
        int rc_ = (call);                                                                             \
        if (rc_ != LUMACU_OK) {                                                                       \
            fprintf(stderr, "%s -> %s: %s\n", #call, lumacu_status_name(rc_), lumacu_last_error(ctx)); \
            return 1;                                                                                 \
        }                                                                                             \
    } while (0)

int main(void)
{
    const uint32_t w = 1280, h = 720;
    const int profile = 2; /* 4:2:0, 16-bit sample containers */
    lumacu_ctx *ctx = NULL;
    CHECK(lumacu_create(0, &ctx));

    float lut[2048];
    CHECK(lumacu_build_lut(LUMACU_PTF_PQ, 11, 1e4f, 0.005f, lut, 2048));
    CHECK(lumacu_set_quantizer(ctx, lut, 2048, 255, LUMACU_CS_LUV, 1e4f));

    const size_t npx = (size_t)w * h;
    float *rgb = NULL, *back = NULL;
    CHECK(lumacu_host_alloc(npx * 3 * sizeof(float), (void **)&rgb)); /* page-locked like the facade's LumaFrame */
    CHECK(lumacu_host_alloc(npx * 3 * sizeof(float), (void **)&back));
    for (size_t i = 0; i < npx; i++) { /* a grey ramp 0.01 .. 5000 cd/m2 with a colour cast */
        const float v = 0.01f * powf(5.0e5f, (float)(i % w) / (float)w);
        rgb[i] = v;
        rgb[npx + i] = 0.8f * v;
        rgb[2 * npx + i] = 0.6f * v;
    }
    const int32_t strides[3] = {(int32_t)(w * 2), (int32_t)w, (int32_t)w};
    uint8_t *planes[3];
    planes[0] = (uint8_t *)malloc((size_t)strides[0] * h);
    planes[1] = (uint8_t *)malloc((size_t)strides[1] * (h / 2));
    planes[2] = (uint8_t *)malloc((size_t)strides[2] * (h / 2));

    lumacu_frame_stats st;
    CHECK(lumacu_encode(ctx, rgb, w, h, profile, 1.0f, planes, strides, 0, &st));
    CHECK(lumacu_decode(ctx, (const uint8_t *const *)planes, strides, w, h, profile, 1.0f, back));

    double worst = 0.0;
    for (size_t i = 0; i < npx; i++) {
        const double e = fabs((double)back[npx + i] - rgb[npx + i]) / rgb[npx + i];
        if (e > worst)
            worst = e;
    }
    printf("mean luminance %.3f cd/m2, max %.1f; worst relative round-trip error of G: %.4f; %llu kernel launches\n",
           st.sum / (double)npx, st.max, worst, (unsigned long long)lumacu_launch_count(ctx));
    lumacu_host_free(rgb);
    lumacu_host_free(back);
    free(planes[0]);
    free(planes[1]);
    free(planes[2]);
    lumacu_destroy(ctx);
    return worst < 0.05 ? 0 : 2;
}
