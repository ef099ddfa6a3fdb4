#include "frt_internal.cuh"
void frt_bank_release(frt_ctx *) {}
