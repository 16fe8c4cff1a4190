// tables.cu — host copy of the packed marching-cubes case tables (uploaded once per ctx).
#include "mc_tables.inc"
const signed char *b2m_mc_table_blob(void) { return MCT_BLOB; }
