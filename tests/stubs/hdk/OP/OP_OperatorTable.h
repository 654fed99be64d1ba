#pragma once  // TEST INFRASTRUCTURE (see SYS/SYS_Types.h)
#include <OP/OP_Operator.h>
class OP_OperatorTable {
   public:
	bool addOperator(OP_Operator* op);
};
