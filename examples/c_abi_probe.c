/* Binds libunirec_b200.so from plain C through the public header: proves that include/unirec_b200.h is a C header
 * (no C++ / torch types), that the library loads without Python, and exercises the host-only entry points.
 *   gcc -std=c11 -Iinclude examples/c_abi_probe.c -ldl -o /tmp/c_abi_probe && /tmp/c_abi_probe unirec_b200/libunirec_b200.so
 * Compute entry points need a GPU and device pointers; they are declared in the header and only looked up here. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>

#include "unirec_b200.h"

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "unirec_b200/libunirec_b200.so";
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "dlopen failed: %s\n", dlerror()); return 2; }
    int (*abi_version)(void) = (int (*)(void))dlsym(h, "unirec_abi_version");
    int64_t (*ws_bytes)(int64_t, int64_t, int64_t) = (int64_t (*)(int64_t, int64_t, int64_t))dlsym(h, "unirec_score_topk_workspace_bytes");
    const char* (*last_error)(void) = (const char* (*)(void))dlsym(h, "unirec_last_error");
    if (!abi_version || !ws_bytes || !last_error) { fprintf(stderr, "missing symbol\n"); return 3; }
    if (abi_version() != UNIREC_B200_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 4; }
    const char* compute[] = {"unirec_linear_bf16", "unirec_attention", "unirec_layernorm", "unirec_score_topk",
                             "unirec_topk_merge", "unirec_build_user_sequence", "unirec_linear_gather_bf16",
                             "unirec_list_scores", "unirec_infonce_rank", "unirec_inject_tokens",
                             "unirec_reconstruction_metrics", "unirec_layernorm_backward_fused"};
    for (unsigned i = 0; i < sizeof(compute) / sizeof(compute[0]); ++i)
        if (!dlsym(h, compute[i])) { fprintf(stderr, "missing %s\n", compute[i]); return 5; }
    int64_t b = ws_bytes(4096, 1000000, 100);
    if (b <= 0) { fprintf(stderr, "workspace query failed: %s\n", last_error()); return 6; }
    /* a NULL-pointer call must come back as an error code with a message, not crash (no GPU needed: arguments are
     * validated before anything is launched) */
    int (*cast)(const float*, void*, int64_t, void*) = (int (*)(const float*, void*, int64_t, void*))dlsym(h, "unirec_cast_f32_to_bf16");
    int rc = cast(NULL, NULL, 0, NULL);
    printf("abi %d, score_topk workspace for 4096 x 1M x top-100: %lld bytes, bad call -> rc %d (%s)\n", abi_version(),
           (long long)b, rc, last_error());
    dlclose(h);
    return rc != 0 ? 0 : 7;
}
