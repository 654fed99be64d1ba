#pragma once  // TEST INFRASTRUCTURE (see SYS/SYS_Types.h): the real header defines the DSO version symbol
