/* Plain-C consumer of include/dyffusion_b200.h (test infrastructure): proves the boundary is a C ABI -- the header compiles
 * as C99 without any C++ / torch type, a C program links against libdyffusion_b200.so, and, on a machine without a GPU,
 * every compute entry point answers with its documented error code and message instead of falling back to the CPU.
 * Built and run by tests/test_c_abi_cpu.py. */
#include <stdio.h>
#include <string.h>

#include "dyffusion_b200.h"

static int fails = 0;
#define EXPECT(cond)                                                      \
  do {                                                                    \
    if (!(cond)) { printf("FAIL line %d: %s\n", __LINE__, #cond); ++fails; } \
  } while (0)

int main(void) {
  dyf_net_desc d;
  dyf_net* net = NULL;
  size_t bytes = 0;
  int n_keys, i, found = 0;
  int64_t first = 0;
  float dummy[16] = {0};

  EXPECT(dyf_abi_version() == DYF_ABI_VERSION);

  /* spring-mesh forecaster (src/models/simple_conv_net.py): host-side handle logic needs no device */
  memset(&d, 0, sizeof d);
  d.arch = DYF_ARCH_CONVNET; d.dim = 64; d.in_channels = 4; d.cond_channels = 5; d.out_channels = 4;
  d.height = 10; d.width = 10; d.with_time_emb = 1; d.dropout = 0.05f;
  d.n_kernels = 4; d.kernel_sizes[0] = 9; d.kernel_sizes[1] = 7; d.kernel_sizes[2] = 5; d.kernel_sizes[3] = 3; d.residual = 1;
  EXPECT(dyf_net_create(&d, &net) == DYF_OK && net != NULL);
  n_keys = dyf_net_num_params(net);
  EXPECT(n_keys > 10);
  for (i = 0; i < n_keys; ++i)
    if (strcmp(dyf_net_param_key(net, i), "time_emb_mlp.1.weight") == 0) found = 1;  /* reference state-dict key */
  EXPECT(found);
  EXPECT(dyf_net_set_param(net, "no.such.key", dummy, &first, 1) < 0 && strlen(dyf_last_error()) > 0);
  EXPECT(dyf_net_finalize(net, NULL) < 0);  /* no device here; on a GPU box: parameters never set, names the first missing key */
  printf("finalize: %s\n", dyf_last_error());
  dyf_net_destroy(net);

  d.arch = 99;
  EXPECT(dyf_net_create(&d, &net) < 0);

  /* argument errors are reported before any device work */
  EXPECT(dyf_window_gather(NULL, 0, 0, NULL, 0, 0, NULL, NULL) == DYF_ERR_ARG);
  EXPECT(dyf_adamw_workspace_bytes(1024, &bytes) == DYF_OK && bytes >= sizeof(double));
  EXPECT(dyf_adamw_step(NULL, NULL, NULL, NULL, 0, 1e-3, 0.9, 0.99, 1e-8, 0.0, 1, 0.0, NULL, 0, NULL) == DYF_ERR_ARG);
  EXPECT(dyf_boundary_conditions_spring_mesh(NULL, NULL, NULL, 1, 1, 10, 10, NULL) == DYF_ERR_ARG);

#ifdef EXPECT_NO_GPU
  /* valid arguments, no device: DYF_ERR_CUDA, never a CPU result */
  first = 0;
  EXPECT(dyf_window_gather(dummy, 4, 4, &first, 1, 2, dummy, NULL) == DYF_ERR_CUDA);
  EXPECT(strstr(dyf_last_error(), "no CPU fallback") != NULL);
  EXPECT(dyf_boundary_conditions_spring_mesh(dummy, (const uint8_t*)dummy, dummy, 1, 1, 1, 1, NULL) == DYF_ERR_CUDA);
#endif
  printf(fails ? "c_abi_probe: %d failure(s)\n" : "c_abi_probe: ok\n", fails);
  return fails;
}
