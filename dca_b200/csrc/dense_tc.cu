// tcgen05 / TMEM / TMA Dense kernels for the gene-wide layers -- placeholder hooks until the
// kernels land (the generic path in dense_generic.cu serves every shape meanwhile).
#include "engine.h"

namespace dca {

Engine::~Engine() { for (auto e : prof.ev) cudaEventDestroy(e); }
bool Engine::tc_supported() const { return false; }
const char* Engine::tc_reason() const { return "tcgen05 kernels not built into this library version"; }
int Engine::setup_tc() { return DCA_OK; }
int Engine::refresh_shadows(cudaStream_t) { return DCA_OK; }

}  // namespace dca
