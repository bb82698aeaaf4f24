// viewer::RenderOptions IS the C-ABI options record (include/mnv_b200.h, mnv_render_options): every field the
// reference's struct has (include/render_options.hpp:9-56) is inherited under its own name, in the same order and
// with the same types; the defaults (:12-55) are filled in by mnv_render_options_default(), the one place they are
// written down.  Code written against the reference compiles unchanged, and the object goes to any C-ABI call as it is.
#pragma once

#include "../../../include/mnv_b200.h"

#define VIEWER_GLOBAL_BASIS_MAX MNV_GLOBAL_BASIS_MAX

namespace viewer {

struct RenderOptions : mnv_render_options {
    RenderOptions() { mnv_render_options_default(this); }
};

}  // namespace viewer
