#include "Backend.hpp"
namespace abl {
void CudaBackend::generate(Script &, const BackendContext &) { throw BackendError("cuda backend: not implemented yet"); }
}
