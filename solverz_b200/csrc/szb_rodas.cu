// szb_rodas.cu -- batched Rodas driver (placeholder until the stage kernels land).
#include "szb.h"
#include "szb_internal.cuh"

extern "C" int szb_rodas(szb_handle* h, const szb_rodas_opt*, double*, double*, int32_t*, szb_stream) {
  szb_set_error(h, "szb_rodas: not built yet");
  return SZB_ESTATE;
}
