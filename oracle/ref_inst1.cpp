/* TEST INFRASTRUCTURE ONLY (oracle/): explicit instantiation of the reference templates for spans 32 and 64,
 * playing the role of the cmake-generated gatb/template/TemplateSpecialization*.cpp.in files (we do not run cmake). */
#include <gatb/kmer/impl/Model.cpp>
#include <gatb/kmer/impl/ConfigurationAlgorithm.cpp>
#include <gatb/kmer/impl/RepartitionAlgorithm.cpp>
namespace gatb { namespace core { namespace kmer { namespace impl  {
#define INST(K) template struct Kmer<K>; template class ConfigurationAlgorithm<K>; template class RepartitorAlgorithm<K>;
INST(32)
INST(64)
}}}}
