/* Plain-C client of libb200env.so: the headers compile as C99, every declared entry point resolves,
 * and the calls that need no GPU behave as documented.  Built and run by tests/test_host_layer.py. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "b200data.h"
#include "b200env.h"
#include "b200policy.h"

#define RESOLVE(name)                                        \
    do {                                                     \
        if (!dlsym(lib, #name)) {                            \
            fprintf(stderr, "missing symbol %s\n", #name);   \
            return 2;                                        \
        }                                                    \
    } while (0)

int main(int argc, char **argv) {
    void *lib;
    int (*abi_version)(void);
    size_t (*minmax_workspace)(int);
    size_t (*rank_workspace)(void);
    int (*column_minmax)(const void *, int, int64_t, int, double *, double *, void *, void *);
    const char *(*data_error)(void);
    int (*policy_create)(int, int, int, b2p_handle *);
    const char *(*policy_error)(b2p_handle);
    b2p_handle policy = NULL;
    if (argc < 2) return 64;
    lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 1; }
    RESOLVE(b2e_abi_version); RESOLVE(b2e_create); RESOLVE(b2e_destroy); RESOLVE(b2e_last_error);
    RESOLVE(b2e_num_params); RESOLVE(b2e_obs_dim); RESOLVE(b2e_history_depth); RESOLVE(b2e_bind_dataset);
    RESOLVE(b2e_set_index_stream); RESOLVE(b2e_reset); RESOLVE(b2e_step); RESOLVE(b2e_eval);
    RESOLVE(b2e_get_state); RESOLVE(b2e_set_state); RESOLVE(b2e_get_batch_indices); RESOLVE(b2e_next_batch);
    RESOLVE(b2e_set_trace); RESOLVE(b2e_get_trace); RESOLVE(b2e_launch_count);
    RESOLVE(b2d_last_error); RESOLVE(b2d_resize_nearest); RESOLVE(b2d_minmax_workspace);
    RESOLVE(b2d_column_minmax); RESOLVE(b2d_normalize); RESOLVE(b2d_rank_workspace);
    RESOLVE(b2d_label_ranks); RESOLVE(b2d_onehot); RESOLVE(b2d_shuffle_permutations);
    RESOLVE(b2p_create); RESOLVE(b2p_destroy); RESOLVE(b2p_last_error); RESOLVE(b2p_set_weights);
    RESOLVE(b2p_act); RESOLVE(b2p_act_env); RESOLVE(b2p_set_seed_counter);
    *(void **)&abi_version = dlsym(lib, "b2e_abi_version");
    *(void **)&minmax_workspace = dlsym(lib, "b2d_minmax_workspace");
    *(void **)&rank_workspace = dlsym(lib, "b2d_rank_workspace");
    *(void **)&column_minmax = dlsym(lib, "b2d_column_minmax");
    *(void **)&data_error = dlsym(lib, "b2d_last_error");
    if (abi_version() != 2) return 3;
    if (minmax_workspace(49) == 0 || minmax_workspace(0) != 0 || rank_workspace() < 2 * 65536 * 4) return 4;
    /* argument errors are reported before any CUDA call */
    if (column_minmax(NULL, B2D_U8, 0, 49, NULL, NULL, NULL, NULL) != B2D_EINVAL) return 5;
    if (!strstr(data_error(), "b2d_column_minmax")) return 6;
    *(void **)&policy_create = dlsym(lib, "b2p_create");
    *(void **)&policy_error = dlsym(lib, "b2p_last_error");
    if (policy_create(0, B2P_MAX_OBS_DIM + 1, B2P_TANH_F32, &policy) == 0 || policy != NULL) return 7;
    if (!strstr(policy_error(NULL), "obs_dim")) return 8;
    printf("c abi ok\n");
    return 0;
}
