// the elementwise kernels of the Float instance at every variant (lbk_tables.hpp)
#include "lbk_tables.hpp"
LBK_MAP_TABLE_DEFINITION
LBK_MAP_TABLES(float)
