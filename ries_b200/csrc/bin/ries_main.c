/* `ries` command: the drop-in for the reference binary (gcc -o ries ries.c -lm). */
#include <stdio.h>
#include "../../../include/ries_b200.h"

int main(int argc, char **argv) {
  int code = ries_main(argc, argv, 0, 0);
  fflush(stdout);
  return code; /* exit() keeps the low 8 bits: -9998 -> 242, -1 -> 255, as with the reference */
}
