// the reductions of the Double instance at every variant (lbk_tables.hpp)
#include "lbk_tables.hpp"
LBK_REDUCE_TABLE_DEFINITION
LBK_REDUCE_TABLES(double)
