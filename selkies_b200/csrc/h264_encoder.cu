// placeholder; replaced by the real encoder
#include "h264_encoder.h"
namespace b2v {
struct Encoder { int dummy; };
int encoder_create(const EncoderConfig*, Encoder**) { return -4; }
void encoder_destroy(Encoder*) {}
size_t encoder_au_capacity(const Encoder*) { return 0; }
int encoder_encode(Encoder*, const EncodeFrameParams*, cudaStream_t) { return 0; }
const uint8_t* encoder_recon(const Encoder*) { return nullptr; }
const char* encoder_last_error() { return "encoder not built yet"; }
}
