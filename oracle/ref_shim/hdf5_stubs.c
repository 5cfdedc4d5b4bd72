/* TEST INFRASTRUCTURE ONLY (oracle/): every HDF5 entry point aborts -- the oracle uses the reference's file storage. */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
int64_t H5Aclose() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Aclose reached\n"); abort(); }
int64_t H5Acreate2() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Acreate2 reached\n"); abort(); }
int64_t H5Adelete() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Adelete reached\n"); abort(); }
int64_t H5Aexists() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Aexists reached\n"); abort(); }
int64_t H5Aget_space() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Aget_space reached\n"); abort(); }
int64_t H5Aopen() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Aopen reached\n"); abort(); }
int64_t H5Aread() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Aread reached\n"); abort(); }
int64_t H5Awrite() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Awrite reached\n"); abort(); }
int64_t H5Dclose() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Dclose reached\n"); abort(); }
int64_t H5Dcreate2() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Dcreate2 reached\n"); abort(); }
int64_t H5Dget_space() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Dget_space reached\n"); abort(); }
int64_t H5Dopen2() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Dopen2 reached\n"); abort(); }
int64_t H5Dread() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Dread reached\n"); abort(); }
int64_t H5Dset_extent() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Dset_extent reached\n"); abort(); }
int64_t H5Dvlen_reclaim() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Dvlen_reclaim reached\n"); abort(); }
int64_t H5Dwrite() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Dwrite reached\n"); abort(); }
int64_t H5Eset_auto() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Eset_auto reached\n"); abort(); }
int64_t H5Fclose() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Fclose reached\n"); abort(); }
int64_t H5Fcreate() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Fcreate reached\n"); abort(); }
int64_t H5Fopen() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Fopen reached\n"); abort(); }
int64_t H5Gclose() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Gclose reached\n"); abort(); }
int64_t H5Gcreate2() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Gcreate2 reached\n"); abort(); }
int64_t H5Gopen2() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Gopen2 reached\n"); abort(); }
int64_t H5Lexists() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Lexists reached\n"); abort(); }
int64_t H5Pclose() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Pclose reached\n"); abort(); }
int64_t H5Pcreate() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Pcreate reached\n"); abort(); }
int64_t H5Pset_chunk() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Pset_chunk reached\n"); abort(); }
int64_t H5Pset_deflate() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Pset_deflate reached\n"); abort(); }
int64_t H5Pset_layout() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Pset_layout reached\n"); abort(); }
int64_t H5Pset_shuffle() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Pset_shuffle reached\n"); abort(); }
int64_t H5Sclose() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Sclose reached\n"); abort(); }
int64_t H5Screate_simple() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Screate_simple reached\n"); abort(); }
int64_t H5Sget_simple_extent_dims() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Sget_simple_extent_dims reached\n"); abort(); }
int64_t H5Sselect_hyperslab() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Sselect_hyperslab reached\n"); abort(); }
int64_t H5Tclose() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Tclose reached\n"); abort(); }
int64_t H5Tcopy() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Tcopy reached\n"); abort(); }
int64_t H5Tcreate() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Tcreate reached\n"); abort(); }
int64_t H5Tinsert() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Tinsert reached\n"); abort(); }
int64_t H5Tpack() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Tpack reached\n"); abort(); }
int64_t H5Tset_precision() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Tset_precision reached\n"); abort(); }
int64_t H5Tset_size() { fprintf(stderr, "oracle/_ref: HDF5 stub H5Tset_size reached\n"); abort(); }
