/* TEST INFRASTRUCTURE ONLY (oracle/): declaration-only stand-in for <hdf5/hdf5.h>.
 *
 * The reference's k-mer headers (gatb/tools/math/LargeInt.hpp:38, gatb/tools/misc/api/Abundance.hpp:34,
 * gatb/tools/storage/impl/StorageHDF5.hpp) include the vendored HDF5 public header, which only exists after the
 * reference's cmake configure step has generated H5pubconf.h.  The oracle build never runs cmake: it compiles the
 * reference's own k-mer counting sources where they lie under /root/reference with g++ and uses the reference's
 * "-storage-type file" sink, so no HDF5 function is ever *called*.  This header therefore only DECLARES the
 * symbols those headers name; hdf5_stubs.c defines them to abort() if one is ever reached.
 */
#pragma once
#include <stdint.h>
#include <stddef.h>
typedef int64_t hid_t; typedef int herr_t; typedef int htri_t;
typedef unsigned long long hsize_t; typedef long long hssize_t;
typedef struct { size_t len; void* p; } hvl_t;
#define HOFFSET(S,M) (offsetof(S,M))
#define H5S_UNLIMITED ((hsize_t)(-1))
#define H5Gcreate H5Gcreate2
#define H5D_CHUNKED ((hid_t)100)
#define H5F_ACC_RDONLY ((hid_t)101)
#define H5F_ACC_RDWR ((hid_t)102)
#define H5F_ACC_TRUNC ((hid_t)103)
#define H5P_DATASET_CREATE ((hid_t)104)
#define H5P_DEFAULT ((hid_t)105)
#define H5S_SELECT_SET ((hid_t)106)
#define H5T_COMPOUND ((hid_t)107)
#define H5T_C_S1 ((hid_t)108)
#define H5T_NATIVE_INT ((hid_t)109)
#define H5T_NATIVE_UINT16 ((hid_t)110)
#define H5T_NATIVE_UINT32 ((hid_t)111)
#define H5T_NATIVE_UINT64 ((hid_t)112)
#define H5T_NATIVE_UINT8 ((hid_t)113)
#define H5T_VARIABLE ((hid_t)114)
#ifdef __cplusplus
extern "C" {
#endif
hid_t H5Aclose(...);
hid_t H5Acreate2(...);
hid_t H5Adelete(...);
hid_t H5Aexists(...);
hid_t H5Aget_space(...);
hid_t H5Aopen(...);
hid_t H5Aread(...);
hid_t H5Awrite(...);
hid_t H5Dclose(...);
hid_t H5Dcreate2(...);
hid_t H5Dget_space(...);
hid_t H5Dopen2(...);
hid_t H5Dread(...);
hid_t H5Dset_extent(...);
hid_t H5Dvlen_reclaim(...);
hid_t H5Dwrite(...);
hid_t H5Eset_auto(...);
hid_t H5Fclose(...);
hid_t H5Fcreate(...);
hid_t H5Fopen(...);
hid_t H5Gclose(...);
hid_t H5Gcreate2(...);
hid_t H5Gopen2(...);
hid_t H5Lexists(...);
hid_t H5Pclose(...);
hid_t H5Pcreate(...);
hid_t H5Pset_chunk(...);
hid_t H5Pset_deflate(...);
hid_t H5Pset_layout(...);
hid_t H5Pset_shuffle(...);
hid_t H5Sclose(...);
hid_t H5Screate_simple(...);
hid_t H5Sget_simple_extent_dims(...);
hid_t H5Sselect_hyperslab(...);
hid_t H5Tclose(...);
hid_t H5Tcopy(...);
hid_t H5Tcreate(...);
hid_t H5Tinsert(...);
hid_t H5Tpack(...);
hid_t H5Tset_precision(...);
hid_t H5Tset_size(...);
#ifdef __cplusplus
}
#endif
