#pragma once
#define INT128_FOUND 1
#define KSIZE_LIST 32,64
#define KSIZE_STRING "32 64"
#define KSIZE_LIST_TYPE boost::mpl::int_<32>,boost::mpl::int_<64>
#define CUSTOM_MEM_ALLOC 0
#define GATB_HDF5_NB_ITEMS_PER_BLOCK (4*1024)
#define GATB_HDF5_CLEANUP_WORKAROUND 4
