// Runtime (energy, simultaneous, exact, damped) -> template instantiation, shared by xf_kernels.cu and xf_batch.cu.
#pragma once

#include "xf_scene.h"

namespace xf {

template <template <int, bool, bool, bool> class Fn, typename... Args>
cudaError_t DispatchConfig(uint32_t energy, bool simul, bool exact, bool damped, Args&&... args) {
#define XF_CASE(E)                                                                                                      \
	case E:                                                                                                             \
		if (simul) {                                                                                                    \
			if (exact) { return damped ? Fn<E, true, true, true>::Run(args...) : Fn<E, true, true, false>::Run(args...); } \
			return damped ? Fn<E, true, false, true>::Run(args...) : Fn<E, true, false, false>::Run(args...);           \
		}                                                                                                               \
		if (exact) { return damped ? Fn<E, false, true, true>::Run(args...) : Fn<E, false, true, false>::Run(args...); } \
		return damped ? Fn<E, false, false, true>::Run(args...) : Fn<E, false, false, false>::Run(args...);
	switch (energy) {
		XF_CASE(XF_ENERGY_MIXED)
		XF_CASE(XF_ENERGY_MIXED_SEL)
		XF_CASE(XF_ENERGY_YEOH_SKIN)
		XF_CASE(XF_ENERGY_YEOH_SKIN_FAST)
	default: return cudaErrorInvalidValue;
	}
#undef XF_CASE
}


}  // namespace xf
