/* Drop-in include path: programs written against the reference's <libhydrium/libhydrium.h>
 * compile unchanged against libhydrium_b200.so. */
#include "../hydrium_b200.h"
