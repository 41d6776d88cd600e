/* TEST INFRASTRUCTURE ONLY: -Wl,--wrap=malloc target for the CPU build of the file-scan harness: the reference reads parts of its malloc()ed
 * 46 MB scratch struct before writing them; zero-initialised memory makes it a pure function of its input. */
#include <stdlib.h>
extern "C" void *__wrap_malloc(size_t n) { return calloc(1, n ? n : 1); }
