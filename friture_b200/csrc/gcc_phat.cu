#include "frt_internal.cuh"
void frt_gcc_release(frt_ctx *) {}
