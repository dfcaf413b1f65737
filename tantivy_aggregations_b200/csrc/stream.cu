// stream.cu — streaming fast shapes (tile-staged kernels).  Placeholder until the first shapes land:
// every plan currently runs on the generic tree-walking kernel.
#include "exec.h"

int stream_try(ExecState& es) {
    (void)es;
    return 0;
}
