/* Minimal stand-in for R's <R.h>, ONLY to compile-check src/scdc_rcall.c where R is not installed. Prototypes follow
 * the public R API ("Writing R Extensions"); nothing here is linked or executed. */
#ifndef R_STUB_R_H
#define R_STUB_R_H
#include <stddef.h>
char* R_alloc(size_t n, int size);
#endif
