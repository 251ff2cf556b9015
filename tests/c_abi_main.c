/* tests/c_abi_main.c -- the C ABI seen from a plain C99 caller (gcc -std=c99 -pedantic -Werror): the header must be valid C
 * (no C++-isms outside its extern "C" guard) and the host-side entry points must be callable without a GPU. */
#include <stdio.h>
#include <string.h>

#include "b200_dmz.h"

int main(void) {
  b200_scanner *s = b200_scanner_new();
  uint8_t digits[16];
  int32_t n = 0;
  int complete;
  if (!s) return 2;
  complete = b200_scanner_result(s, digits, &n); /* nothing added: not complete */
  b200_scanner_reset(s);
  b200_scanner_free(s);
  printf("%d %d %d %d %d %d\n", complete, (int)n, (int)sizeof(b200_frame_record), (int)sizeof(b200_scan), (int)sizeof(b200_edges),
         (int)sizeof(b200_expiry_group));
  return 0;
}
