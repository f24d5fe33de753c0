// flow_emu.cpp -- pyro2_b200/csrc/flow.cu (kernels and host orchestration, unchanged) compiled for the
// host through cuda_emu.h.  TEST INFRASTRUCTURE ONLY: built by tests/test_flow_emulated.py into
// tests/emu/libflow_emu.so, which exports the p2b_flow_* C ABI over HOST memory so that the Burgers /
// incompressible stage kernels can be checked bit-for-bit against the oracle without a GPU.
#include <stdarg.h>

#include "cuda_emu_runtime.inc"

#include "../../pyro2_b200/csrc/flow.cu"

namespace pyro {
char* last_error_buf()
{
    static thread_local char buf[512];
    return buf;
}
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
}
}  // namespace pyro

extern "C" const char* p2b_last_error(void) { return pyro::last_error_buf(); }

namespace {
struct RegisterThreaded {
    RegisterThreaded() { emu::threaded((const void*)pyro::flow_maxabs_kernel); }
} register_threaded;
}  // namespace
