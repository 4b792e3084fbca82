/*
 * A plain C host driving one DDIM + SCG step through the C ABI (include/rgm_b200.h): no Python, no torch types.
 * This is the call sequence of INTEGRATION.md section C with every buffer a raw device pointer.  It is compiled as C99
 * by tests/test_capi_cpu.py (the header must be valid C and every entry point must link); on a B200 it runs one step on
 * synthetic weights and prints the chosen candidate indices.
 *
 *   gcc -std=c99 -Wall -I include examples/scg_step_host.c -L rule_guided_music_b200 -lrgm_b200 \
 *       -L /usr/local/cuda/lib64 -lcudart -lm -o scg_step_host
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rgm_b200.h"

/* the few CUDA runtime calls a host needs, declared here so the example does not depend on cuda_runtime.h being C-clean */
extern int cudaMalloc(void** p, size_t n);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind);
extern int cudaMemset(void* p, int v, size_t n);
extern int cudaDeviceSynchronize(void);
enum { H2D = 1, D2H = 2 };

#define CHECK(call)                                              \
  do {                                                           \
    if ((call) != 0) {                                           \
      fprintf(stderr, "%s failed: %s\n", #call, rgm_last_error()); \
      return 1;                                                  \
    }                                                            \
  } while (0)

static unsigned lcg_state = 12345u;
static float lcg_normalish(void) { /* sum of 4 uniforms: enough for synthetic weights / noise */
  float s = 0.f;
  for (int i = 0; i < 4; ++i) {
    lcg_state = lcg_state * 1664525u + 1013904223u;
    s += (float)(lcg_state >> 8) / 16777216.0f - 0.5f;
  }
  return s * 1.7320508f;
}
static float* dev_random(size_t n, float scale) {
  float* h = (float*)malloc(n * sizeof(float));
  float* d = NULL;
  for (size_t i = 0; i < n; ++i) h[i] = lcg_normalish() * scale;
  if (cudaMalloc((void**)&d, n * sizeof(float)) != 0) return NULL;
  cudaMemcpy(d, h, n * sizeof(float), H2D);
  free(h);
  return d;
}
static int load_dit(rgm_dit* dit, const char* key, size_t n, float scale) {
  float* d = dev_random(n, scale);
  int rc = d ? rgm_dit_load(dit, key, d, (long long)n, NULL) : -1;
  cudaDeviceSynchronize();
  return rc < 0 ? rc : 0;
}

int main(void) {
  enum { B = 2, N = 4, C = 4, H = 128, W = 16, T = 256, DEPTH = 2, D = 1152, HEADS = 16, MLP = 4608 };
  const long long elems = (long long)C * H * W;
  if (rgm_check_device() != 0) {
    fprintf(stderr, "no B200: %s\n", rgm_last_error());
    return 2;
  }
  /* schedule tables of the 256-step respacing (respace.py:63-95): betas_i = 1 - abar_i / abar_prev over the kept steps */
  double betas[T], base[1000], abar = 1.0, last = 1.0;
  for (int i = 0; i < 1000; ++i) base[i] = 1e-4 + (2e-2 - 1e-4) * i / 999.0;
  for (int i = 0, k = 0; i < 1000; ++i) {
    abar *= 1.0 - base[i];
    if (k < T && (i == (int)floor(k * 999.0 / (T - 1) + 0.5))) { /* evenly spaced kept steps (illustration) */
      betas[k++] = 1.0 - abar / last;
      last = abar;
    }
  }
  float* tab = NULL;
  cudaMalloc((void**)&tab, sizeof(float) * RGM_COEF_ROWS * T);
  CHECK(rgm_coeff_tables(betas, T, tab, NULL));

  rgm_dit* dit = NULL;
  rgm_vae* vae = NULL;
  rgm_scg* scg = NULL;
  CHECK(rgm_dit_create(&dit, DEPTH, D, HEADS, 8, C, C, 4, W, MLP));
  char key[128];
  CHECK(load_dit(dit, "x_embedder.MLP.0.weight", 256 * 32, 0.1f));
  CHECK(load_dit(dit, "x_embedder.MLP.2.weight", (size_t)D * 256, 0.05f));
  CHECK(load_dit(dit, "t_embedder.mlp.0.weight", (size_t)D * 256, 0.05f));
  CHECK(load_dit(dit, "t_embedder.mlp.2.weight", (size_t)D * D, 0.02f));
  CHECK(load_dit(dit, "final_layer.linear.weight", 32 * (size_t)D, 0.02f));
  CHECK(load_dit(dit, "final_layer.adaLN_modulation.1.weight", 2 * (size_t)D * D, 0.02f));
  for (int i = 0; i < DEPTH; ++i) {
    snprintf(key, sizeof key, "blocks.%d.attn.qkv.weight", i);
    CHECK(load_dit(dit, key, 3 * (size_t)D * D, 0.02f));
    snprintf(key, sizeof key, "blocks.%d.attn.proj.weight", i);
    CHECK(load_dit(dit, key, (size_t)D * D, 0.02f));
    snprintf(key, sizeof key, "blocks.%d.mlp.fc1.weight", i);
    CHECK(load_dit(dit, key, (size_t)MLP * D, 0.02f));
    snprintf(key, sizeof key, "blocks.%d.mlp.fc2.weight", i);
    CHECK(load_dit(dit, key, (size_t)D * MLP, 0.02f));
    snprintf(key, sizeof key, "blocks.%d.adaLN_modulation.1.weight", i);
    CHECK(load_dit(dit, key, 6 * (size_t)D * D, 0.02f));
  }
  {
    float f[128], r[18];
    for (int i = 0; i < 128; ++i) f[i] = expf(-logf(10000.f) * i / 128.f);   /* dit.py:57-59 */
    for (int i = 0; i < 18; ++i) r[i] = 1.f / powf(10000.f, 2.f * i / 36.f); /* RotaryEmbedding(36).freqs */
    float *df, *dr;
    cudaMalloc((void**)&df, sizeof f);
    cudaMalloc((void**)&dr, sizeof r);
    cudaMemcpy(df, f, sizeof f, H2D);
    cudaMemcpy(dr, r, sizeof r, H2D);
    CHECK(rgm_dit_load(dit, "__timestep_freqs", df, 128, NULL));
    CHECK(rgm_dit_load(dit, "rotary_emb.freqs", dr, 18, NULL));
  }
  const int mult[4] = {1, 2, 2, 4};
  CHECK(rgm_vae_create(&vae, 128, mult, 4, 2, 4, 3)); /* zero weights decode to a constant roll: enough to run the step */
  CHECK(rgm_scg_create(&scg, dit, vae));
  CHECK(rgm_scg_reserve(scg, N, B, C, H, W));

  float* x = dev_random((size_t)B * elems, 1.f);
  float* noise = dev_random((size_t)N * B * elems, 1.f);
  float *eps, *x0, *mean, *sigma, *t_model, *target, *scores;
  long long *t_index, *y, *chosen;
  cudaMalloc((void**)&eps, sizeof(float) * B * elems);
  cudaMalloc((void**)&x0, sizeof(float) * B * elems);
  cudaMalloc((void**)&mean, sizeof(float) * B * elems);
  cudaMalloc((void**)&sigma, sizeof(float) * B);
  cudaMalloc((void**)&t_model, sizeof(float) * B);
  cudaMalloc((void**)&target, sizeof(float) * B * 12);
  cudaMalloc((void**)&scores, sizeof(float) * N * B);
  cudaMalloc((void**)&t_index, sizeof(long long) * B);
  cudaMalloc((void**)&y, sizeof(long long) * B);
  cudaMalloc((void**)&chosen, sizeof(long long) * B);
  const int step = 200;
  float ht[B], htarget[B * 12];
  long long hi[B], hy[B], hchosen[B];
  memset(htarget, 0, sizeof htarget);
  for (int b = 0; b < B; ++b) {
    ht[b] = (float)(step * 999 / (T - 1));
    hi[b] = step;
    hy[b] = 1;
    htarget[b * 12] = 0.5f;
    htarget[b * 12 + 4] = 0.25f;
    htarget[b * 12 + 7] = 0.25f;
  }
  cudaMemcpy(t_model, ht, sizeof ht, H2D);
  cudaMemcpy(t_index, hi, sizeof hi, H2D);
  cudaMemcpy(y, hy, sizeof hy, H2D);
  cudaMemcpy(target, htarget, sizeof htarget, H2D);

  /* p_mean_variance + the DDIM algebra, then the SCG step */
  CHECK(rgm_dit_forward(dit, x, t_model, y, eps, B, H, NULL));
  CHECK(rgm_ddim_mean(x, eps, tab, T, t_index, 1.f, 1, x0, mean, sigma, B, elems, NULL));
  float *a, *c;
  cudaMalloc((void**)&a, sizeof(float) * B);
  cudaMalloc((void**)&c, sizeof(float) * B);
  for (int b = 0; b < B; ++b) {
    cudaMemcpy(a + b, tab + (size_t)RGM_COEF_SQRT_RECIP_ALPHAS_CUMPROD * T + step, sizeof(float), 3 /* D2D */);
    cudaMemcpy(c + b, tab + (size_t)RGM_COEF_SQRT_RECIPM1_ALPHAS_CUMPROD * T + step, sizeof(float), 3);
  }
  rgm_rule_spec rule = {RGM_RULE_PITCH_HIST, 128, 5.f, 0, 1.f, target};
  CHECK(rgm_scg_step(scg, mean, sigma, noise, t_model, y, a, c, 1.2465f, &rule, 1, N, B, C, H, W, x, chosen, scores, NULL));
  if (cudaDeviceSynchronize() != 0) {
    fprintf(stderr, "device error\n");
    return 1;
  }
  cudaMemcpy(hchosen, chosen, sizeof hchosen, D2H);
  printf("scg step ok: chosen candidates");
  for (int b = 0; b < B; ++b) printf(" %lld", hchosen[b]);
  printf(" (of %d), %llu kernels launched\n", N, rgm_launch_count());
  rgm_scg_destroy(scg);
  rgm_vae_destroy(vae);
  rgm_dit_destroy(dit);
  return 0;
}
