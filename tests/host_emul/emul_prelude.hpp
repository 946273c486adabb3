// emul_prelude.hpp - force-included (-include) in every product source of the full emulation build
// (tests/host_emul/build_full_emul.sh): the CTA emulator in place of the CUDA execution model.
#pragma once
#define CHB_HOST_EMUL 1
#define CHB_HOST_EMUL_FULL 1
#include "cta_emul.hpp"
