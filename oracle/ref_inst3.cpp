/* TEST INFRASTRUCTURE ONLY (oracle/): explicit instantiation of the reference templates for spans 32 and 64,
 * playing the role of the cmake-generated gatb/template/TemplateSpecialization*.cpp.in files (we do not run cmake). */
#include <gatb/kmer/impl/BloomAlgorithm.cpp>
namespace gatb { namespace core { namespace kmer { namespace impl  {
template class BloomAlgorithm<32>; template class BloomAlgorithm<64>;
}}}}
