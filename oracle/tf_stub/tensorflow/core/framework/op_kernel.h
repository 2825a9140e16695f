// Stand-in header (test infrastructure) — see stub_core.h.
#include "tensorflow/core/framework/stub_core.h"
