/* abi_smoke.c -- a C caller of libb381.so compiled against include/b381.h with gcc: what a cgo binding sees.
 * Catches what ctypes cannot: a prototype that does not match its definition at link/ABI level (argument widths, struct
 * layouts and padding, enum values), or a declared function that the library does not export.
 *
 *   abi_smoke --layout            sizes / offsets / enum values of the header, every declared function referenced
 *                                 (tests/c_abi/symbols.inc is generated from the header by the test); b381_init without
 *                                 a device must return B381_ERR_NO_DEVICE -- there is no CPU fallback
 *   abi_smoke in.bin out.bin      GPU: reads n, then n G1 affine and n G2 affine PODs; writes n Fq12 (b381_pairing_batch),
 *                                 then the verdicts of two b381_pairing_product_is_one groups built from pair 0:
 *                                 {(P0, Q0), (-P0... supplied as P[n-1]) , Q0)} and {(P0, Q0), (P1, Q0)}
 */
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b381.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "abi_smoke: %s failed (line %d)\n", #c, __LINE__); return 1; } } while (0)

typedef void (*anyfn)(void);
static const anyfn exported[] = {
#include "symbols.inc"
};

static int layout(void) {
    CHECK(sizeof(b381_fp) == 48 && sizeof(b381_fp2) == 96 && sizeof(b381_fp12) == 576);
    CHECK(sizeof(b381_g1_affine) == 104 && offsetof(b381_g1_affine, y) == 48 && offsetof(b381_g1_affine, infinity) == 96);
    CHECK(sizeof(b381_g2_affine) == 200 && offsetof(b381_g2_affine, y) == 96 && offsetof(b381_g2_affine, infinity) == 192);
    CHECK(sizeof(b381_g1_jac) == 144 && sizeof(b381_g2_jac) == 288 && sizeof(b381_scalar) == 32);
    CHECK(B381_OK == 0 && B381_ERR_ARG == -1 && B381_ERR_CUDA == -2 && B381_ERR_NOMEM == -3 && B381_ERR_NO_DEVICE == -4);
    CHECK(B381_PATH_AUTO == -1 && B381_PATH_THREAD == 0 && B381_PATH_VM == 1 && B381_PATH_QUAD == 2 && B381_PATH_DUO == 3);
    CHECK(B381_POINT_OK == 0 && B381_POINT_ERR_MODE == 1 && B381_POINT_ERR_INFINITY == 2 && B381_POINT_ERR_NOT_ON_CURVE == 3);
    size_t nsym = sizeof exported / sizeof exported[0];
    for (size_t i = 0; i < nsym; i++) CHECK(exported[i] != NULL);
    CHECK(b381_init(0, NULL) == B381_ERR_ARG);
    CHECK(b381_launch_count(NULL) == 0);
    CHECK(strcmp(b381_last_error(NULL), "null ctx") == 0);
    b381_ctx *ctx = (b381_ctx *)1;
    int rc = b381_init(0, &ctx);
    if (rc == B381_ERR_NO_DEVICE) { CHECK(ctx == NULL); printf("layout ok, %zu exports, no device: B381_ERR_NO_DEVICE\n", nsym); return 0; }
    CHECK(rc == B381_OK && ctx != NULL);
    CHECK(b381_set_kernel_path(ctx, 7) == B381_ERR_ARG && b381_set_kernel_path(ctx, B381_PATH_AUTO) == B381_OK);
    CHECK(b381_pairing_batch(ctx, NULL, NULL, 0, NULL) == B381_OK);
    CHECK(b381_pairing_batch(ctx, NULL, NULL, 1, NULL) == B381_ERR_ARG);
    b381_free(ctx);
    printf("layout ok, %zu exports, device present\n", nsym);
    return 0;
}

int main(int argc, char **argv) {
    if (argc == 2 && strcmp(argv[1], "--layout") == 0) return layout();
    if (argc != 3) { fprintf(stderr, "usage: abi_smoke --layout | abi_smoke in.bin out.bin\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    CHECK(f != NULL);
    uint64_t n = 0;
    CHECK(fread(&n, 8, 1, f) == 1 && n >= 3 && n < (1u << 20));
    b381_g1_affine *P = malloc(n * sizeof *P);
    b381_g2_affine *Q = malloc(n * sizeof *Q);
    b381_fp12 *out = malloc(n * sizeof *out);
    CHECK(P && Q && out);
    CHECK(fread(P, sizeof *P, n, f) == n && fread(Q, sizeof *Q, n, f) == n);
    fclose(f);
    b381_ctx *ctx = NULL;
    CHECK(b381_init(0, &ctx) == B381_OK);
    CHECK(b381_pairing_batch(ctx, P, Q, n, out) == B381_OK);
    /* group 0: e(P0, Q0) e(P[n-1], Q0) with P[n-1] = -P0 -> 1;  group 1: e(P0, Q0) e(P1, Q0) -> not 1 */
    b381_g1_affine gp[4] = {P[0], P[n - 1], P[0], P[1]};
    b381_g2_affine gq[4] = {Q[0], Q[0], Q[0], Q[0]};
    uint32_t off[3] = {0, 2, 4};
    uint8_t ok[2] = {9, 9};
    CHECK(b381_pairing_product_is_one(ctx, gp, gq, 4, off, 2, ok) == B381_OK);
    uint32_t bad_off[3] = {0, 3, 2};
    CHECK(b381_pairing_product_is_one(ctx, gp, gq, 4, bad_off, 2, ok + 0) == B381_ERR_ARG);
    uint64_t launches = b381_launch_count(ctx);
    b381_free(ctx);
    f = fopen(argv[2], "wb");
    CHECK(f != NULL);
    CHECK(fwrite(out, sizeof *out, n, f) == n && fwrite(ok, 1, 2, f) == 2 && fwrite(&launches, 8, 1, f) == 1);
    fclose(f);
    printf("pairings %llu launches %llu verdicts %u %u\n", (unsigned long long)n, (unsigned long long)launches, ok[0], ok[1]);
    return 0;
}
