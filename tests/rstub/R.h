/* see Rinternals.h in this directory */
#include <stddef.h>
#include <stdint.h>
