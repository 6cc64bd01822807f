/* Plain-C caller of the derl_b200 C ABI (no Python, no torch): what a non-Python binding sees.
 * Runs derl_b200_gae_host on a small rollout and compares with the same recursion written
 * inline (reference arithmetic: derl/runners/trajectory_transforms.py:45-63).
 * exit 0 = bit-identical, 2 = no sm_100 device (DERL_E_NO_DEVICE), 1 = mismatch/other error. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "derl_b200.h"

int main(void) {
  enum { T = 37, N = 48 };
  static double rewards[T * N];
  static float values[T * N], last_value[N], adv[T * N], vt[T * N], want[T * N];
  static uint8_t resets[T * N];
  unsigned s = 12345u;
  for (int i = 0; i < T * N; ++i) {
    s = s * 1664525u + 1013904223u;
    rewards[i] = ((int)(s >> 8) % 2001 - 1000) / 250.0;
    s = s * 1664525u + 1013904223u;
    values[i] = (float)(((int)(s >> 8) % 2001 - 1000) / 333.0);
    s = s * 1664525u + 1013904223u;
    resets[i] = (s >> 16) % 10 == 0;
  }
  for (int n = 0; n < N; ++n) last_value[n] = (float)(n - 20) / 7.0f;
  const double gamma = 0.99, lambda = 0.95;
  if (derl_b200_abi_version() != DERL_B200_ABI_VERSION) return 1;
  int rc = derl_b200_gae_host(rewards, 1, values, resets, last_value, T, N, gamma, lambda, 0, 1e-8,
                              adv, vt, NULL);
  if (rc == DERL_E_NO_DEVICE) {
    printf("no device: %s\n", derl_b200_last_error());
    return 2;
  }
  if (rc != DERL_OK) {
    printf("error %d: %s\n", rc, derl_b200_last_error());
    return 1;
  }
  for (int n = 0; n < N; ++n) {
    int i = (T - 1) * N + n;
    float base = (float)(rewards[i] - (double)values[i]);
    double nr = resets[i] ? 0.0 : 1.0;
    want[i] = (float)((double)base + (nr * gamma) * (double)last_value[n]);
    for (int t = T - 2; t >= 0; --t) {
      i = t * N + n;
      nr = resets[i] ? 0.0 : 1.0;
      double delta = (rewards[i] + (nr * gamma) * (double)values[i + N]) - (double)values[i];
      want[i] = (float)(delta + ((nr * gamma) * lambda) * (double)want[i + N]);
    }
  }
  for (int i = 0; i < T * N; ++i) {
    float target = want[i] + values[i];
    if (memcmp(&adv[i], &want[i], 4) != 0 || memcmp(&vt[i], &target, 4) != 0) {
      printf("mismatch at %d: %.9g vs %.9g\n", i, adv[i], want[i]);
      return 1;
    }
  }
  printf("abi smoke ok: %d elements bit-identical, %llu kernel launches\n", T * N,
         (unsigned long long)derl_b200_launch_count());
  return 0;
}
