// forwards to the serial stand-in (oracle/ref_shim/libmesh/shim.h); test infrastructure only
#include "shim.h"
