/* TEST INFRASTRUCTURE ONLY (oracle/): explicit instantiation of the reference templates for spans 32 and 64,
 * playing the role of the cmake-generated gatb/template/TemplateSpecialization*.cpp.in files (we do not run cmake). */
#include <gatb/kmer/impl/SortingCountAlgorithm.cpp>
#include <gatb/kmer/impl/PartitionsCommand.cpp>
namespace gatb { namespace core { namespace kmer { namespace impl  {
#define INST(K) template class SortingCountAlgorithm<K>; template class PartitionsCommand<K>; template class PartitionsByHashCommand<K>; template class PartitionsByVectorCommand<K>; 
INST(32)
INST(64)
}}}}
