#pragma once
#define STR_LIBRARY_VERSION     "1.4.2"
#define STR_COMPILATION_DATE    "oracle-direct-build"
#define STR_COMPILATION_FLAGS   "-std=c++11 -O3 -DNDEBUG"
#define STR_COMPILER            "g++"
#define STR_OPERATING_SYSTEM    "Linux"
