/* TEST DOUBLE (oracle only): forwards to the loopback libvpx stand-in. */
#include "vpx_stub.h"
