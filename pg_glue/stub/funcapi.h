/* stub: everything lives in postgres.h of this directory (pg_glue/stub/README) */
#include "postgres.h"
