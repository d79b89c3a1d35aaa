/* bgpt_tables.c -- host-side builder of ggml's two fp16 lookup tables (ggml.c:4620-4640):
 *   table_gelu_f16[i] = fp16(gelu(f32(i)))      ggml_gelu_f32, ggml.c:3842-3844
 *   table_exp_f16[i]  = fp16(expf(f32(i)))
 * Built with the host libm at model-load time, exactly as the reference builds them in
 * ggml_init, so that the device looks up the very values the reference would on this machine.
 * Compiled with -ffp-contract=off: the one fused multiply-add the reference binary contains in
 * its GELU (gcc -O3 -mfma contracts `1 + a*x*x`) is spelled fmaf() here.
 */
#include <immintrin.h>
#include <math.h>
#include <stdint.h>

static float gelu_tanh_f32(float x) {
    const float coef_a = 0.044715f;
    const float sqrt_2_over_pi = 0.79788456080286535587989211986876f;
    const float poly = fmaf(coef_a * x, x, 1.0f);
    return (0.5f * x) * (1.0f + tanhf((sqrt_2_over_pi * x) * poly));
}

void bgpt_host_build_tables(uint16_t * gelu_f16, uint16_t * exp_f16) {
    for (int i = 0; i < 65536; i++) {
        const float f = _cvtsh_ss((uint16_t) i);
        gelu_f16[i] = _cvtss_sh(gelu_tanh_f32(f), 0);
        exp_f16[i]  = _cvtss_sh(expf(f), 0);
    }
}
