// Temporary: entry points not built yet report DME_EINVAL instead of being absent.
#include "type_quantize.cuh"
namespace dme {
int biased_quantize(const float *, int64_t, int64_t, int64_t, int64_t, const WsLayout &, void *, int32_t *, uint8_t *, float *, int64_t,
                    uint32_t *, int64_t, uint64_t *, cudaStream_t) { set_error("biased mode: not built yet"); return DME_EINVAL; }
}
#define NOTYET(name) { dme::set_error(name ": not built yet"); return DME_EINVAL; }
extern "C" {
int dme_drive(const float *, int64_t, int64_t, int64_t, float *, int64_t, uint64_t, const float *, int, dme_stream_t) NOTYET("dme_drive")
int dme_eden_encode(const float *, int64_t, int64_t, int64_t, int64_t, int, uint64_t, const float *, const float *, float *, uint8_t *, float *, dme_stream_t) NOTYET("dme_eden_encode")
int dme_eden_decode(const uint8_t *, const float *, int64_t, int64_t, int64_t, int, uint64_t, const float *, float *, float *, int64_t, dme_stream_t) NOTYET("dme_eden_decode")
int dme_quicfl_decode(const int32_t *, const int32_t *, int64_t, int64_t, int64_t, int, const float *, int, const uint8_t *, const float *, const int64_t *, const float *, uint64_t, const float *, float *, float *, int64_t, dme_stream_t) NOTYET("dme_quicfl_decode")
int dme_scalar_quantize(const float *, int64_t, int64_t, int64_t, float, uint64_t, uint64_t, const float *, float *, int64_t, dme_stream_t) NOTYET("dme_scalar_quantize")
}
