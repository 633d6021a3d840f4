/* A C99 host of include/rbl.h: proves the header is plain C (no C++ leaks across the ABI) and
 * that a program linked against librbl.so alone can drive the path.  Without a GPU it checks the
 * documented failure mode (rbl_create fails loudly, no CPU fallback); with one it runs a tiny
 * product.  Built and run by tests/test_capi_and_shells.py. */
#include <stdio.h>
#include <string.h>

#include "rbl.h"

int main(void) {
  rbl_ctx* ctx = NULL;
  int st = rbl_create(RBL_F64, -1, &ctx);
  if (st != RBL_OK) {
    const char* msg = rbl_last_error(NULL);
    printf("NO-DEVICE status=%d msg=%s\n", st, msg ? msg : "(null)");
    return (msg && strstr(msg, "no CPU fallback")) ? 0 : 2;
  }
  double ref[3] = {0, 0, 0};
  double r[6] = {0, 0, 1.0, 0.5, 0, 1.2}, F[6] = {1, 0, 0, 0, 1, 0}, U[6];
  if (rbl_set_parameters(ctx, 0.1, 0.01, 1.0, 1.0, ref, 1) != RBL_OK) return 3;
  if (rbl_set_flags(ctx, 0, 1) != RBL_OK) return 4;
  if (rbl_apply_M(ctx, F, r, 2, U) != RBL_OK) { printf("%s\n", rbl_last_error(ctx)); return 5; }
  printf("DEVICE-OK %s U0=%.6f\n", rbl_version(), U[0]);
  rbl_destroy(ctx);
  return 0;
}
