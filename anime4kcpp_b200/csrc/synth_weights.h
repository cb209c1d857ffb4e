/*
 * Deterministic synthetic weights for the ARNet<8> family.
 *
 * The reference's trained ARNet numbers live in
 * core/internal/AC/Core/Internal/Model/Param/ARNet.p, which is listed in the
 * reference's .MISSING_LARGE_BLOBS and is absent from the tree this build was
 * made against.  The ARNet *architecture* is fully specified
 * (core/include/AC/Core/Model/ARNet.hpp:16-72), so every ARNet variant is given
 * seeded stand-in weights with exactly the reference's array lengths.  The
 * generator uses integer hashing and exact float arithmetic only (no libm), so
 * the product library, the CPU oracle and the compiled-reference stand-in all
 * obtain bit-identical arrays.  Real weights can be supplied at run time through
 * acb200_model_create_custom().
 */
#ifndef AC_B200_SYNTH_WEIGHTS_H
#define AC_B200_SYNTH_WEIGHTS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

static inline uint64_t acsw_splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* uniform in [0,1) with 24 random bits: exactly representable in fp32 */
static inline float acsw_uniform(uint64_t *s)
{
    return (float)(acsw_splitmix64(s) >> 40) * (1.0f / 16777216.0f);
}

/* Irwin-Hall(4) approximation of N(0,1): sum of 4 uniforms has variance 1/3 */
static inline float acsw_normal(uint64_t *s)
{
    float u = acsw_uniform(s);
    u += acsw_uniform(s);
    u += acsw_uniform(s);
    u += acsw_uniform(s);
    return (u - 2.0f) * 1.7320508f;
}

static inline uint64_t acsw_seed_from_name(const char *name)
{
    uint64_t h = 0xCBF29CE484222325ULL; /* FNV-1a */
    while (*name) { h ^= (uint8_t)*name++; h *= 0x100000001B3ULL; }
    return h;
}

/* ARNet<8> array lengths, core/include/AC/Core/Model/ARNet.hpp:32-37 */
static inline int acsw_arnet_kernel_len(int blocks) { return 8 * 9 + 8 * 8 * 9 * blocks * 2 + 8 * 8 + 8 * 4 * 9; }
static inline int acsw_arnet_bias_len(int blocks) { return 8 + 8 * (blocks * 2 + 1) + 4; }
static inline int acsw_arnet_alpha_len(int blocks) { return 8 * (blocks + 1); }

/*
 * Fill the three flat arrays of an ARNet<8> variant.  `name` is the canonical
 * model string ("arnet-f8b16-box-hdn", ...); it only seeds the generator.
 */
static inline void acsw_fill_arnet(const char *name, int blocks, float *kernels, float *biases, float *alphas)
{
    uint64_t s = acsw_seed_from_name(name);
    int nk = acsw_arnet_kernel_len(blocks);
    int nb = acsw_arnet_bias_len(blocks);
    int na = acsw_arnet_alpha_len(blocks);
    int body_end = 72 + 576 * blocks * 2;
    int i;
    for (i = 0; i < nk; i++)
    {
        float sigma = (i < 72) ? 0.2f : (i < body_end ? 0.05f : (i < body_end + 64 ? 0.2f : 0.05f));
        kernels[i] = acsw_normal(&s) * sigma;
    }
    for (i = 0; i < nb; i++) biases[i] = acsw_normal(&s) * 0.02f;
    for (i = 0; i < na; i++) alphas[i] = 0.05f + 0.25f * acsw_uniform(&s);
}

#ifdef __cplusplus
}
#endif

#endif
