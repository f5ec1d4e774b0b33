// the reductions of the Float instance (and sdsdot) at every variant (lbk_tables.hpp)
#include "lbk_tables.hpp"
LBK_REDUCE_TABLE_DEFINITION
LBK_REDUCE_TABLES(float)
template const void *reduce_kernel_ptr<lbk::DsdotOp>(int);
