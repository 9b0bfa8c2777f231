/* TEST DOUBLE (see mkv_interface_stub.h): lets code that includes "mkv_interface.h" by name build
 * against the in-memory container when the reference's include directory is not on the path. */
#include "mkv_interface_stub.h"
