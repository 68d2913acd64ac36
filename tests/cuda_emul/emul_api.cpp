// emul_api.cpp - TEST INFRASTRUCTURE ONLY: switches of the CPU emulation exported next to the C ABI in libljgpu_emul.so.
#include "cuda_emul.h"

// Shared-memory gather profile: on = 1 starts counting from zero, on = 0 stops; the three outputs (any may be null)
// receive the warp-level 64-bit gather instructions, their wavefronts and their active lanes since the start.
extern "C" void ljemul_lds_profile(int on, unsigned long long* instr, unsigned long long* wavefronts, unsigned long long* lanes) {
    if (instr) *instr = ljemul::g_lds.instr.load();
    if (wavefronts) *wavefronts = ljemul::g_lds.wavefronts.load();
    if (lanes) *lanes = ljemul::g_lds.lanes.load();
    if (on) { ljemul::g_lds.instr = 0; ljemul::g_lds.wavefronts = 0; ljemul::g_lds.lanes = 0; }
    ljemul::g_lds.on = on != 0;
}
