/*
 * TEST INFRASTRUCTURE.  Emits a stand-in for the reference's missing weight file
 * core/internal/AC/Core/Internal/Model/Param/ARNet.p (listed in .MISSING_LARGE_BLOBS):
 * the 16 array triples core/src/Model.cpp:129-224 binds, filled with the seeded synthetic
 * numbers of anime4kcpp_b200/csrc/synth_weights.h (printed with %.9g, which round-trips
 * fp32 exactly).  Used only by oracle/build_ref.sh so the compiled reference runs ARNet on
 * the same weights as the product and the oracle.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../anime4kcpp_b200/csrc/synth_weights.h"

static void emit(const char *sym, const char *what, const float *v, int n)
{
    printf("alignas(AC_CORE_PARAM_ALIGN) constexpr float %s_NHWC_%s[] = {\n", sym, what);
    for (int i = 0; i < n; i++) printf("%s%.9gf,%s", (i % 8 == 0) ? "  " : " ", v[i], (i % 8 == 7 || i == n - 1) ? "\n" : "");
    printf("};\n");
}

int main(void)
{
    static const int blocks[4] = { 8, 16, 32, 64 };
    static const char *vsym[4] = { "", "_HDN", "_Box", "_Box_HDN" };
    static const char *vname[4] = { "", "-hdn", "-box", "-box-hdn" };
    for (int bi = 0; bi < 4; bi++)
        for (int vi = 0; vi < 4; vi++)
        {
            int B = blocks[bi];
            char name[64], sym[64];
            snprintf(name, sizeof(name), "arnet-f8b%d%s", B, vname[vi]);
            snprintf(sym, sizeof(sym), "ARNet_F8B%d%s", B, vsym[vi]);
            int nk = acsw_arnet_kernel_len(B), nb = acsw_arnet_bias_len(B), na = acsw_arnet_alpha_len(B);
            float *k = malloc(sizeof(float) * nk), *b = malloc(sizeof(float) * nb), *a = malloc(sizeof(float) * na);
            acsw_fill_arnet(name, B, k, b, a);
            emit(sym, "kernels", k, nk);
            emit(sym, "biases", b, nb);
            emit(sym, "alphas", a, na);
            free(k); free(b); free(a);
        }
    return 0;
}
