// ghost_emu.cpp -- pyro2_b200/csrc/ghost_cfl.cu (ghost-cell fill and CFL wave-speed kernels with their host entry
// points, unchanged) compiled for the host through cuda_emu.h.  TEST INFRASTRUCTURE ONLY: built by
// tests/emu_util.py into tests/emu/libghost_emu.so, which exports p2b_fill_ghost_*, p2b_cfl_wavemax and the error
// plumbing over HOST memory; the product only ever loads the nvcc-built libpyro2b200.so.
#include "cuda_emu_runtime.inc"

#include "../../pyro2_b200/csrc/ghost_cfl.cu"
#include "../../pyro2_b200/csrc/slab_comm.cu"

namespace {
struct RegisterThreaded {
    RegisterThreaded() { emu::threaded((const void*)pyro::cfl_kernel); }
} register_threaded;
}  // namespace
