// gatb_core_b200.hpp -- host-side C++ mirror of GATB-core's API for the k-mer counting path, above the C ABI
// (include/gatb_gpu.h).  Header-only; same names, argument meaning and error behaviour as the reference so that code
// written against GATB-core compiles against this for THIS path:
//
//   Kmer<span>::{Type, ModelDirect, ModelCanonical, ModelMinimizer<ModelCanonical>, Count}   kmer/impl/Model.hpp:89-1592
//   Configuration                                                                          kmer/impl/Configuration.hpp:40-121
//   Repartitor (operator(), load/save of the "minimRepart" byte stream)                   kmer/impl/PartiInfo.hpp:292-387, PartiInfo.cpp:228-295
//   ICountProcessor<span>, CountProcessorAbstract/Chain/Histogram/SolidityInfo/Dump        kmer/api/ICountProcessor.hpp:91-183, kmer/impl/CountProcessor*.hpp
//   SortingCountAlgorithm<span>                                                            kmer/impl/SortingCountAlgorithm.hpp:65-263
//   IBank / Sequence / BankStrings / BankFasta (minimal)                                   bank/api/IBank.hpp:78-148
//   Histogram (inc, compute_threshold)                                                     tools/misc/impl/Histogram.hpp:92, Histogram.cpp:61-190
//   BloomBuilder (sizing + build on the device)                                            kmer/impl/BloomBuilder.hpp:102-131, BloomAlgorithm.cpp:155-203
//
// The k-mer models below are plain host arithmetic (they are the API downstream tools use to interpret k-mers); the
// counting itself is NOT re-implemented on the host: SortingCountAlgorithm<span>::execute() forwards to gatb_gpu_count
// and throws system::Exception when no CUDA device is available (no CPU fallback).
#ifndef GATB_CORE_B200_HPP
#define GATB_CORE_B200_HPP

#include "../../../include/gatb_gpu.h"

#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace gatb { namespace core {

/********************************************************************************/
namespace system {
/** system/api/Exception.hpp:59 */
class Exception
{
public:
    Exception () {}
    Exception (const char* format, ...)
    {
        char buf[1024]; va_list ap; va_start (ap, format); vsnprintf (buf, sizeof(buf), format, ap); va_end (ap);
        _message = buf;
    }
    const char* getMessage () const { return _message.c_str(); }
protected:
    std::string _message;
};
/** system/api/ISmartPointer.hpp:97-130 (intrusive reference counting: use / forget) */
class SmartPointer
{
public:
    SmartPointer () : _ref(0) {}
    virtual ~SmartPointer () {}
    void use    () { ++_ref; }
    void forget () { if (--_ref <= 0) delete this; }
private:
    int _ref;
};
template<class T> inline void setAttr (T*& attr, T* value) { if (value) value->use(); if (attr) attr->forget(); attr = value; }
} // system

/********************************************************************************/
namespace tools { namespace math {
/** tools/math/LargeInt1.pri (u64) and LargeInt2.pri (__uint128_t): the two integer types of the path */
template<int N> class LargeInt;
template<> class LargeInt<1>
{
public:
    typedef uint64_t raw;
    LargeInt (uint64_t v = 0) : value(v) {}
    static const char* getName () { return "LargeInt<1>"; }
    static size_t getSize () { return 64; }
    uint64_t getVal () const { return value; }
    void setVal (uint64_t v) { value = v; }
    uint64_t lo () const { return value; }
    uint64_t hi () const { return 0; }
    static LargeInt make (uint64_t lo, uint64_t) { return LargeInt (lo); }
    raw value;
};
template<> class LargeInt<2>
{
public:
    typedef unsigned __int128 raw;
    LargeInt (uint64_t v = 0) : value(v) {}
    static const char* getName () { return "LargeInt<2>"; }
    static size_t getSize () { return 128; }
    uint64_t getVal () const { return (uint64_t)value; }
    void setVal (uint64_t v) { value = v; }
    uint64_t lo () const { return (uint64_t)value; }
    uint64_t hi () const { return (uint64_t)(value >> 64); }
    static LargeInt make (uint64_t lo, uint64_t hi) { LargeInt r; r.value = ((raw)hi << 64) | lo; return r; }
    raw value;
};
#define GATB_B200_OP(op) template<int N> inline LargeInt<N> operator op (const LargeInt<N>& a, const LargeInt<N>& b) { LargeInt<N> r; r.value = a.value op b.value; return r; }
GATB_B200_OP(+) GATB_B200_OP(-) GATB_B200_OP(&) GATB_B200_OP(|) GATB_B200_OP(^)
#undef GATB_B200_OP
template<int N> inline LargeInt<N> operator<< (const LargeInt<N>& a, int s) { LargeInt<N> r; r.value = a.value << s; return r; }
template<int N> inline LargeInt<N> operator>> (const LargeInt<N>& a, int s) { LargeInt<N> r; r.value = a.value >> s; return r; }
template<int N> inline bool operator<  (const LargeInt<N>& a, const LargeInt<N>& b) { return a.value <  b.value; }
template<int N> inline bool operator<= (const LargeInt<N>& a, const LargeInt<N>& b) { return a.value <= b.value; }
template<int N> inline bool operator== (const LargeInt<N>& a, const LargeInt<N>& b) { return a.value == b.value; }
template<int N> inline bool operator!= (const LargeInt<N>& a, const LargeInt<N>& b) { return a.value != b.value; }

/** tools/math/LargeInt1.pri:137-154 */
inline uint64_t revcomp64 (uint64_t x, size_t sizeKmer)
{
    if (sizeKmer == 0) return 0;
    uint64_t res = x;
    res = ((res >>  2) & 0x3333333333333333ULL) | ((res & 0x3333333333333333ULL) <<  2);
    res = ((res >>  4) & 0x0F0F0F0F0F0F0F0FULL) | ((res & 0x0F0F0F0F0F0F0F0FULL) <<  4);
    res = ((res >>  8) & 0x00FF00FF00FF00FFULL) | ((res & 0x00FF00FF00FF00FFULL) <<  8);
    res = ((res >> 16) & 0x0000FFFF0000FFFFULL) | ((res & 0x0000FFFF0000FFFFULL) << 16);
    res = ((res >> 32) & 0x00000000FFFFFFFFULL) | ((res & 0x00000000FFFFFFFFULL) << 32);
    res ^= 0xAAAAAAAAAAAAAAAAULL;
    return res >> (2 * (32 - sizeKmer));
}
inline LargeInt<1> revcomp (const LargeInt<1>& x, size_t sizeKmer) { return LargeInt<1> (revcomp64 (x.value, sizeKmer)); }
/** tools/math/LargeInt2.pri:168-197 */
inline LargeInt<2> revcomp (const LargeInt<2>& in, size_t sizeKmer)
{
    uint64_t high = in.hi (), low = in.lo ();
    size_t nb_high = sizeKmer > 32 ? sizeKmer - 32 : 0, nb_low = sizeKmer > 32 ? 32 : sizeKmer;
    uint64_t rh = sizeKmer <= 32 ? 0 : revcomp64 (high, nb_high);
    uint64_t rl = revcomp64 (low, nb_low);
    LargeInt<2> res; res.value = rl; res.value <<= 2 * nb_high; res.value += rh;
    return res;
}
/** tools/math/LargeInt1.pri:157-170 */
inline uint64_t hash64 (uint64_t key, uint64_t seed)
{
    uint64_t hash = seed;
    hash ^= (hash <<  7) ^  key * (hash >> 3) ^ (~((hash << 11) + (key ^ (hash >> 5))));
    hash = (~hash) + (hash << 21);
    hash = hash ^ (hash >> 24);
    hash = (hash + (hash << 3)) + (hash << 8);
    hash = hash ^ (hash >> 14);
    hash = (hash + (hash << 2)) + (hash << 4);
    hash = hash ^ (hash >> 28);
    hash = hash + (hash << 31);
    return hash;
}
inline uint64_t hash1 (const LargeInt<1>& k, uint64_t seed = 0) { return hash64 (k.value, seed); }
inline uint64_t hash1 (const LargeInt<2>& k, uint64_t seed = 0) { return hash64 (k.hi (), seed) ^ hash64 (k.lo (), seed); }
template<int N> inline std::string toString (const LargeInt<N>& v, size_t sizeKmer)
{
    static const char bin2NT[4] = {'A','C','T','G'};
    std::string s (sizeKmer, 'A');
    typename LargeInt<N>::raw t = v.value;
    for (size_t i = 0; i < sizeKmer; i++) { s[sizeKmer - 1 - i] = bin2NT[(int)(t & 3)]; t >>= 2; }
    return s;
}
}} // tools::math

/********************************************************************************/
namespace tools { namespace misc {
/** tools/misc/api/Range.hpp:75 (CountRange::includes) */
struct CountRange
{
    CountRange (int64_t b = 0, int64_t e = 0) : begin(b), end(e) {}
    int64_t getBegin () const { return begin; } int64_t getEnd () const { return end; }
    bool includes (int64_t v) const { return begin <= v && v <= end; }
    int64_t begin, end;
};
/** tools/misc/api/Abundance.hpp:68-131 */
template<typename Type, typename Number = int32_t> struct Abundance
{
    Abundance () : abundance(0) {}
    Abundance (const Type& v, const Number& a) : value(v), abundance(a) {}
    const Type& getValue () const { return value; } const Number& getAbundance () const { return abundance; }
    Type value; Number abundance;
};
/** tools/misc/impl/Histogram.hpp / Histogram.cpp:61-190 (the cutoff arithmetic is done by gatb_gpu_histogram_cutoff) */
class Histogram
{
public:
    Histogram (size_t length) : _length(length), _table(length + 1, 0), _cutoff(0), _nbsolids(0), _firstPeak(0) {}
    void inc (uint32_t index) { _table[index >= _length ? _length : index]++; }
    void set (const uint64_t* table) { std::copy (table, table + _length + 1, _table.begin()); }
    uint64_t& get (uint32_t idx) { return _table[idx]; }
    size_t getLength () const { return _length; }
    void compute_threshold (int min_auto_threshold)
    { gatb_gpu_histogram_cutoff (_table.data(), (int)_length, min_auto_threshold, &_cutoff, &_nbsolids, &_firstPeak); }
    uint32_t get_solid_cutoff () const { return _cutoff; } uint64_t get_nbsolids_auto () const { return _nbsolids; }
    uint32_t get_first_peak () const { return _firstPeak; }
private:
    size_t _length; std::vector<uint64_t> _table; uint32_t _cutoff; uint64_t _nbsolids; uint32_t _firstPeak;
};
/** a minimal IProperties: the info tree of Algorithm::getInfo() flattened to key -> value */
class Properties
{
public:
    void add (const std::string& key, const std::string& value) { _m[key] = value; }
    void add (const std::string& key, uint64_t v) { std::stringstream s; s << v; _m[key] = s.str(); }
    bool has (const std::string& key) const { return _m.count (key) != 0; }
    std::string getStr (const std::string& key) const { std::map<std::string,std::string>::const_iterator it = _m.find (key); return it == _m.end() ? "" : it->second; }
    int64_t getInt (const std::string& key) const { return atoll (getStr (key).c_str()); }
    const std::map<std::string,std::string>& all () const { return _m; }
private:
    std::map<std::string,std::string> _m;
};
}} // tools::misc

/********************************************************************************/
namespace bank {
/** bank/api/Sequence.hpp: the nucleotides of one read (ASCII) */
struct Sequence
{
    std::string comment, data;
    const char* getDataBuffer () const { return data.data(); } size_t getDataSize () const { return data.size(); }
};
/** bank/api/IBank.hpp:78-148 reduced to what the counting path consumes: an ordered pass over the sequences */
class IBank : public system::SmartPointer
{
public:
    virtual ~IBank () {}
    virtual std::string getId () = 0;
    virtual void iterate (void (*fct)(const Sequence&, void*), void* arg) = 0;
    virtual int64_t getNbItems () { return -1; }
};
/** bank/impl/BankStrings.hpp */
class BankStrings : public IBank
{
public:
    BankStrings () {}
    BankStrings (const char* s, ...) { va_list ap; va_start (ap, s); for (const char* p = s; p; p = va_arg (ap, const char*)) _seqs.push_back (p); va_end (ap); }
    BankStrings (const std::vector<std::string>& v) : _seqs(v) {}
    std::string getId () { return "strings"; }
    void iterate (void (*fct)(const Sequence&, void*), void* arg) { Sequence s; for (size_t i = 0; i < _seqs.size(); i++) { s.data = _seqs[i]; fct (s, arg); } }
    int64_t getNbItems () { return (int64_t)_seqs.size(); }
private:
    std::vector<std::string> _seqs;
};
/** bank/impl/BankFasta.hpp:65 -- FASTA / FASTQ text files (multi-line FASTA supported; no gzip) */
class BankFasta : public IBank
{
public:
    BankFasta (const std::string& path) : _path(path) {}
    std::string getId () { return _path; }
    void iterate (void (*fct)(const Sequence&, void*), void* arg)
    {
        std::ifstream in (_path.c_str());
        if (!in) throw system::Exception ("Unable to open file '%s'", _path.c_str());
        std::string line; Sequence s; bool have = false; int fastq_state = 0;
        while (std::getline (in, line))
        {
            if (!line.empty() && line[line.size()-1] == '\r') line.erase (line.size()-1);
            if (fastq_state == 1) { s.data = line; fastq_state = 2; continue; }
            if (fastq_state == 2) { fastq_state = 3; continue; }           // '+'
            if (fastq_state == 3) { fct (s, arg); fastq_state = 0; continue; }   // qualities
            if (line.empty()) continue;
            if (line[0] == '>') { if (have) fct (s, arg); s.comment = line.substr (1); s.data.clear(); have = true; }
            else if (line[0] == '@' && !have) { s.comment = line.substr (1); fastq_state = 1; }
            else s.data += line;
        }
        if (have) fct (s, arg);
    }
private:
    std::string _path;
};
} // bank

/********************************************************************************/
namespace kmer {
typedef int32_t CountNumber;                       /* system/api/types.hpp:49 */
typedef std::vector<CountNumber> CountVector;
enum Strand { STRAND_FORWARD = 1, STRAND_REVCOMP = 2 };

namespace impl {
static const unsigned char comp_NT[4] = {2, 3, 0, 1};          /* kmer/impl/ModelData.cpp:41 */
inline bool validNucleotide (unsigned char c) { return c=='A'||c=='C'||c=='G'||c=='T'||c=='a'||c=='c'||c=='g'||c=='t'; }   /* tools/misc/api/Data.cpp:3 */

/** kmer/impl/Model.hpp:89 */
template<size_t span = 32> struct Kmer
{
    typedef tools::math::LargeInt<(span + 31) / 32> Type;                  /* Model.hpp:100 */

    /** Model.hpp:113-196 */
    class KmerDirect
    {
    public:
        const Type& value () const { return _value; }
        const Type& value (int) const { return _value; }
        const Type& forward () const { return _value; }
        bool isValid () const { return _isValid; }
        void set (const Type& v) { _value = v; }
        Type _value; bool _isValid;
    };
    /** Model.hpp:218-323 */
    class KmerCanonical
    {
    public:
        const Type& value () const { return table[(int)choice]; }
        const Type& value (int which) const { return table[which]; }
        const Type& forward () const { return table[0]; }
        const Type& revcomp () const { return table[1]; }
        bool which () const { return choice == 0; }
        Strand strand () const { return which () ? STRAND_FORWARD : STRAND_REVCOMP; }
        bool isValid () const { return _isValid; }
        bool isPalindrome () const { return table[0] == table[1]; }
        void set (const Type& v) { table[0] = v; table[1] = v; choice = 0; }
        void set (const Type& f, const Type& r) { table[0] = f; table[1] = r; updateChoice (); }
        void updateChoice () { choice = (table[0] < table[1]) ? 0 : 1; }
        Type table[2]; char choice; bool _isValid;
    };
    /** Model.hpp:330-374 */
    template<class ModelKmer> class KmerMinimizerT : public ModelKmer
    {
    public:
        const ModelKmer& minimizer () const { return _minimizer; }
        int position () const { return _position; }
        bool hasChanged () const { return _changed; }
        ModelKmer _minimizer; int16_t _position; bool _changed;
    };

    /** Model.hpp:385-770 (ModelAbstract), ModelDirect :780-830, ModelCanonical :840-895 */
    template<class Impl, class K> class ModelAbstract
    {
    public:
        typedef K Kmer;
        ModelAbstract (size_t sizeKmer = span - 1) : _kmerSize(sizeKmer)
        {
            if (sizeKmer >= span) throw system::Exception ("Type '%s' has too low precision (%d bits) for the required %d kmer size", Type::getName(), (int)Type::getSize(), (int)sizeKmer);
            Type un (1); _kmerMask = (un << (int)(_kmerSize * 2)) - un;
            for (int i = 0; i < 4; i++) _revcompTable[i] = Type (comp_NT[i]) << (int)(2 * (_kmerSize - 1));
        }
        size_t getSpan () const { return span; }
        size_t getKmerSize () const { return _kmerSize; }
        const Type& getKmerMax () const { return _kmerMask; }
        std::string toString (const Type& kmer) const { return tools::math::toString (kmer, _kmerSize); }
        Type reverse (const Type& kmer) const { return tools::math::revcomp (kmer, _kmerSize); }
        /** Model.hpp:636-657: returns -1 or the index of the last bad character */
        int polynom (const char* seq, Type& kmer, size_t startIndex) const
        {
            int bad = -1; kmer = Type (0);
            for (size_t i = 0; i < _kmerSize; i++) { unsigned char c = seq[i + startIndex]; kmer = (kmer << 2) + Type ((c >> 1) & 3); if (!validNucleotide (c)) bad = (int)i; }
            return bad;
        }
        Kmer codeSeed (const char* seq, size_t startIndex = 0) const { Kmer r; static_cast<const Impl*>(this)->first (seq, r, startIndex); return r; }
        Kmer codeSeedRight (const Kmer& kmer, char nucl) const { Kmer r = kmer; static_cast<const Impl*>(this)->next ((nucl >> 1) & 3, r, validNucleotide (nucl)); return r; }
        Kmer getKmer (const std::string& data, size_t startIndex = 0) const { return codeSeed (data.data(), startIndex); }
        /** Model.hpp:725-765 */
        template<typename Callback> bool iterate (const char* seq, size_t length, Callback callback) const
        {
            int32_t nbKmers = (int32_t)length - (int32_t)_kmerSize + 1;
            if (nbKmers <= 0) return false;
            Kmer result;
            int indexBadChar = static_cast<const Impl*>(this)->first (seq, result, 0);
            size_t idxComputed = 0;
            callback (result, idxComputed);
            for (size_t idx = _kmerSize; idx < length; idx++)
            {
                unsigned char c = seq[idx];
                if (!validNucleotide (c)) indexBadChar = (int)_kmerSize - 1; else indexBadChar--;
                static_cast<const Impl*>(this)->next ((c >> 1) & 3, result, indexBadChar < 0);
                callback (result, ++idxComputed);
            }
            return true;
        }
        template<typename Callback> bool iterate (const std::string& data, Callback callback) const { return iterate (data.data(), data.size(), callback); }
        bool build (const std::string& data, std::vector<Kmer>& kmersBuffer) const
        { kmersBuffer.clear(); return iterate (data, Pusher (kmersBuffer)); }
    protected:
        struct Pusher { std::vector<Kmer>& v; Pusher (std::vector<Kmer>& v) : v(v) {} void operator() (const Kmer& k, size_t) { v.push_back (k); } };
        size_t _kmerSize; Type _kmerMask; Type _revcompTable[4];
    };
    class ModelDirect : public ModelAbstract<ModelDirect, KmerDirect>
    {
    public:
        typedef KmerDirect Kmer;
        ModelDirect (size_t kmerSize = span - 1) : ModelAbstract<ModelDirect, KmerDirect> (kmerSize) {}
        int first (const char* seq, Kmer& value, size_t startIndex) const { int r = this->polynom (seq, value._value, startIndex); value._isValid = r < 0; return r; }
        void next (char c, Kmer& value, bool isValid) const { value._value = ((value._value << 2) + Type ((uint64_t)c)) & this->_kmerMask; value._isValid = isValid; }
    };
    class ModelCanonical : public ModelAbstract<ModelCanonical, KmerCanonical>
    {
    public:
        typedef KmerCanonical Kmer;
        ModelCanonical (size_t kmerSize = span - 1) : ModelAbstract<ModelCanonical, KmerCanonical> (kmerSize) {}
        int first (const char* seq, Kmer& value, size_t startIndex) const
        {
            int r = this->polynom (seq, value.table[0], startIndex);
            value._isValid = r < 0; value.table[1] = this->reverse (value.table[0]); value.updateChoice ();
            return r;
        }
        void next (char c, Kmer& value, bool isValid) const
        {
            value.table[0] = ((value.table[0] << 2) + Type ((uint64_t)c)) & this->_kmerMask;
            value.table[1] = ((value.table[1] >> 2) + this->_revcompTable[(int)c]) & this->_kmerMask;
            value._isValid = isValid; value.updateChoice ();
        }
        uint64_t getHash (const Type& k) const { return tools::math::hash1 (k, 0); }
    };
    /** Model.hpp:989-1330 (lexicographic order with the mmer_lut; the frequency comparator is not part of this path) */
    template<class ModelType> class ModelMinimizer : public ModelAbstract<ModelMinimizer<ModelType>, KmerMinimizerT<typename ModelType::Kmer> >
    {
    public:
        typedef ModelType Model;
        typedef KmerMinimizerT<typename ModelType::Kmer> Kmer;
        ModelMinimizer (size_t kmerSize, size_t minimizerSize)
            : ModelAbstract<ModelMinimizer<ModelType>, Kmer> (kmerSize), _kmerModel(kmerSize), _miniModel(minimizerSize), _minimizerSize(minimizerSize)
        {
            if (kmerSize < minimizerSize) throw system::Exception ("Bad values for kmer %d and minimizer %d", (int)kmerSize, (int)minimizerSize);
            _nbMinimizers = kmerSize - minimizerSize + 1;
            _mask = ((uint64_t)1 << (2 * _minimizerSize)) - 1;
            uint64_t nb = (uint64_t)1 << (2 * _minimizerSize);
            _mmer_lut.resize (nb);
            for (uint64_t ii = 0; ii < nb; ii++)
            {
                uint64_t mmer = ii, rev = tools::math::revcomp64 (ii, minimizerSize);
                if (rev < mmer) mmer = rev;
                if (!is_allowed ((uint32_t)mmer, (uint32_t)minimizerSize)) mmer = _mask;
                _mmer_lut[ii] = (uint32_t)mmer;
            }
        }
        const ModelType& getMmersModel () const { return _miniModel; }
        int first (const char* seq, Kmer& kmer, size_t startIndex) const
        { int r = _kmerModel.first (seq, kmer, startIndex); computeNewMinimizer (kmer); return r; }
        void next (char c, Kmer& kmer, bool isValid) const
        {
            _kmerModel.next (c, kmer, isValid);
            kmer._isValid = isValid;
            uint64_t mmer = _mmer_lut[kmer.value (0).getVal () & _mask];
            kmer._position--; kmer._changed = false;
            if (mmer < kmer._minimizer.value ().getVal ()) { kmer._minimizer.set (Type (mmer)); kmer._position = (int16_t)(_nbMinimizers - 1); kmer._changed = true; }
            else if (kmer._position < 0) computeNewMinimizer (kmer);
        }
        uint64_t getMinimizerValue (const Type& k) const { Kmer km; km.set (k); computeNewMinimizer (km); return km.minimizer ().value ().getVal (); }
        std::string getMinimizerString (const Type& k) const { Kmer km; km.set (k); computeNewMinimizer (km); return _miniModel.toString (km.minimizer ().value ()); }
    private:
        ModelType _kmerModel, _miniModel; size_t _minimizerSize, _nbMinimizers; uint64_t _mask; std::vector<uint32_t> _mmer_lut;
        /** Model.hpp:1220-1251 */
        static bool is_allowed (uint32_t mmer, uint32_t len)
        {
            uint64_t mask_ma1 = 0x5555555555555555ULL & (((uint64_t)1 << ((len - 2) * 2)) - 1);
            uint64_t a1 = mmer; a1 = ~(a1 | (a1 >> 2)); a1 = ((a1 >> 1) & a1) & mask_ma1;
            return a1 == 0;
        }
        /** Model.hpp:1254-1287 */
        void computeNewMinimizer (Kmer& kmer) const
        {
            kmer._minimizer.set (Type (_mask)); kmer._position = -1; kmer._changed = true;
            uint64_t best = _mask; Type val = kmer.value (0);
            for (int16_t idx = (int16_t)_nbMinimizers - 1; idx >= 0; idx--)
            {
                uint64_t cand = _mmer_lut[val.getVal () & _mask];
                if (cand < best) { kmer._minimizer.set (Type (cand)); kmer._position = idx; best = cand; }
                val = val >> 2;
            }
        }
    };
    /** Model.hpp:1568-1590 */
    struct Count : tools::misc::Abundance<Type, CountNumber>
    {
        Count () {}
        Count (const Type& val, const CountNumber& abund) : tools::misc::Abundance<Type, CountNumber> (val, abund) {}
        bool operator< (const Count& other) const { return this->value < other.value; }
        bool operator== (const Count& other) const { return this->value == other.value && this->abundance == other.abundance; }
    };
};

/** kmer/impl/Configuration.hpp:40-121: the fields that decide the output of the counting path */
struct Configuration
{
    Configuration () : _isComputed(false), _kmerSize(31), _minim_size(10), _repartitionType(0), _minimizerType(0), _max_disk_space(0),
        _max_memory(5000), _nbCores(0), _nb_partitions_in_parallel(0), _partitionType(0), _abundanceUserNb(1),
        _estimateSeqNb(0), _estimateSeqTotalSize(0), _estimateSeqMaxSize(0), _available_space(0), _volume(0), _kmersNb(0),
        _nb_passes(1), _nb_partitions(1), _nb_bits_per_kmer(0), _nb_banks(1), _nb_cached_items_per_core_per_part(0), _histogramMax(10000)
    { _abundance.push_back (tools::misc::CountRange (2, 0x7fffffff)); }
    bool _isComputed;
    size_t _kmerSize, _minim_size, _repartitionType, _minimizerType;
    std::vector<tools::misc::CountRange> _abundance;
    uint64_t _max_disk_space; uint32_t _max_memory; size_t _nbCores, _nb_partitions_in_parallel, _partitionType, _abundanceUserNb;
    uint64_t _estimateSeqNb, _estimateSeqTotalSize, _estimateSeqMaxSize, _available_space, _volume, _kmersNb;
    uint32_t _nb_passes, _nb_partitions; uint16_t _nb_bits_per_kmer, _nb_banks; size_t _nb_cached_items_per_core_per_part;
    size_t _histogramMax;                 /* -histo-max */
};

/** kmer/impl/PartiInfo.hpp:292-387; byte stream of save/load: PartiInfo.cpp:228-295 */
class Repartitor : public system::SmartPointer
{
public:
    typedef uint16_t Value;
    Repartitor (int nbpart = 0, int minimsize = 0, int nbPass = 1) : _nbpart(nbpart), _nb_minims((uint64_t)1 << (minimsize * 2)), _nbPass(nbPass) {}
    Value operator() (uint64_t minimizerValue) const { return _repart_table[minimizerValue]; }
    size_t getNbPasses () const { return _nbPass; }
    uint16_t getNbPartitions () const { return _nbpart; }
    std::vector<Value>& getTable () { return _repart_table; }
    const std::vector<Value>& getTable () const { return _repart_table; }
    void setTable (const std::vector<Value>& t) { _repart_table = t; _nb_minims = t.size(); }
    /** the reference's "minimizers/minimRepart" byte stream: u16 nbpart, u64 nb_minims, u16 nbPass, u16 table[], bool hasFreq, u32 magic */
    void load (std::istream& is)
    {
        bool hasFreq = false; uint32_t magic = 0;
        is.read ((char*)&_nbpart, sizeof(_nbpart)); is.read ((char*)&_nb_minims, sizeof(_nb_minims)); is.read ((char*)&_nbPass, sizeof(_nbPass));
        _repart_table.resize (_nb_minims);
        is.read ((char*)_repart_table.data(), sizeof(Value) * _nb_minims);
        is.read ((char*)&hasFreq, sizeof(bool)); is.read ((char*)&magic, sizeof(magic));
        if (magic != MAGIC_NUMBER) throw system::Exception ("Unable to load Repartitor (minimRepart), possibly due to bad format.");
        if (hasFreq) throw system::Exception ("Repartitor: minimizer frequencies (minimizer type 1) are not supported on the device path");
    }
    void save (std::ostream& os) const
    {
        bool hasFreq = false; uint32_t magic = MAGIC_NUMBER;
        os.write ((const char*)&_nbpart, sizeof(_nbpart)); os.write ((const char*)&_nb_minims, sizeof(_nb_minims)); os.write ((const char*)&_nbPass, sizeof(_nbPass));
        os.write ((const char*)_repart_table.data(), sizeof(Value) * _nb_minims);
        os.write ((const char*)&hasFreq, sizeof(bool)); os.write ((const char*)&magic, sizeof(magic));
    }
    static const uint32_t MAGIC_NUMBER = 0x12345678;
private:
    uint16_t _nbpart; uint64_t _nb_minims; uint16_t _nbPass; std::vector<Value> _repart_table;
};

/********************************************************************************/
/** kmer/api/ICountProcessor.hpp:91-183 -- THE plugin interface results are delivered through */
template<size_t span> class ICountProcessor : public system::SmartPointer
{
public:
    typedef typename Kmer<span>::Type Type;
    virtual ~ICountProcessor () {}
    virtual void begin (const Configuration& config) = 0;
    virtual void end () = 0;
    virtual void beginPass (size_t passId) = 0;
    virtual void endPass (size_t passId) = 0;
    virtual ICountProcessor* clone () = 0;
    virtual void finishClones (std::vector<ICountProcessor<span>*>& clones) = 0;
    virtual void beginPart (size_t passId, size_t partId, size_t cacheSize, const char* name) = 0;
    virtual void endPart (size_t passId, size_t partId) = 0;
    virtual bool process (size_t partId, const Type& kmer, const CountVector& count, CountNumber sum = 0) = 0;
    virtual std::string getName () const = 0;
    virtual std::vector<ICountProcessor*> getInstances () const = 0;
    template<typename T> T* get () const
    { std::vector<ICountProcessor*> v = this->getInstances (); for (size_t i = 0; i < v.size(); i++) if (T* o = dynamic_cast<T*> (v[i])) return o; return (T*)0; }
};
/** kmer/impl/CountProcessorAbstract.hpp */
template<size_t span> class CountProcessorAbstract : public ICountProcessor<span>
{
public:
    typedef typename Kmer<span>::Type Type;
    CountProcessorAbstract (const std::string& name = "processor") : _name(name) {}
    void begin (const Configuration&) {} void end () {} void beginPass (size_t) {} void endPass (size_t) {}
    void finishClones (std::vector<ICountProcessor<span>*>&) {}
    void beginPart (size_t, size_t, size_t, const char*) {} void endPart (size_t, size_t) {}
    bool process (size_t, const Type&, const CountVector&, CountNumber = 0) { return true; }
    std::string getName () const { return _name; }
    std::vector<ICountProcessor<span>*> getInstances () const { std::vector<ICountProcessor<span>*> r; r.push_back ((ICountProcessor<span>*)this); return r; }
protected:
    CountNumber computeSum (const CountVector& count) const { CountNumber s = 0; for (size_t i = 0; i < count.size(); i++) s += count[i]; return s; }
    std::string _name;
};
/** kmer/impl/CountProcessorChain.hpp:128-134 (short-circuit on the first false) */
template<size_t span> class CountProcessorChain : public CountProcessorAbstract<span>
{
public:
    typedef ICountProcessor<span> CountProcessor; typedef typename Kmer<span>::Type Type;
    CountProcessorChain (const std::vector<CountProcessor*>& items) : CountProcessorAbstract<span>("chain"), _items(items) { for (size_t i = 0; i < _items.size(); i++) _items[i]->use (); }
    CountProcessorChain (CountProcessor* a, CountProcessor* b = 0, CountProcessor* c = 0) : CountProcessorAbstract<span>("chain")
    { CountProcessor* v[3] = {a, b, c}; for (int i = 0; i < 3; i++) if (v[i]) { v[i]->use (); _items.push_back (v[i]); } }
    ~CountProcessorChain () { for (size_t i = 0; i < _items.size(); i++) _items[i]->forget (); }
    void begin (const Configuration& c) { for (size_t i = 0; i < _items.size(); i++) _items[i]->begin (c); }
    void end () { for (size_t i = 0; i < _items.size(); i++) _items[i]->end (); }
    void beginPass (size_t p) { for (size_t i = 0; i < _items.size(); i++) _items[i]->beginPass (p); }
    void endPass (size_t p) { for (size_t i = 0; i < _items.size(); i++) _items[i]->endPass (p); }
    CountProcessor* clone () { std::vector<CountProcessor*> c; for (size_t i = 0; i < _items.size(); i++) c.push_back (_items[i]->clone ()); return new CountProcessorChain (c); }
    void finishClones (std::vector<CountProcessor*>& clones)
    {
        for (size_t i = 0; i < _items.size(); i++)
        {
            std::vector<CountProcessor*> sub;
            for (size_t j = 0; j < clones.size(); j++) if (CountProcessorChain* c = dynamic_cast<CountProcessorChain*> (clones[j])) sub.push_back (c->_items[i]);
            _items[i]->finishClones (sub);
        }
    }
    void beginPart (size_t a, size_t b, size_t c, const char* n) { for (size_t i = 0; i < _items.size(); i++) _items[i]->beginPart (a, b, c, n); }
    void endPart (size_t a, size_t b) { for (size_t i = 0; i < _items.size(); i++) _items[i]->endPart (a, b); }
    bool process (size_t partId, const Type& kmer, const CountVector& count, CountNumber sum = 0)
    {
        if (sum == 0) sum = this->computeSum (count);
        bool res = true;
        for (size_t i = 0; res && i < _items.size(); i++) res = _items[i]->process (partId, kmer, count, sum);
        return res;
    }
    std::vector<CountProcessor*> getInstances () const
    { std::vector<CountProcessor*> r; for (size_t i = 0; i < _items.size(); i++) { std::vector<CountProcessor*> v = _items[i]->getInstances (); r.insert (r.end(), v.begin(), v.end()); } return r; }
    const std::vector<CountProcessor*>& items () const { return _items; }
private:
    std::vector<CountProcessor*> _items;
};
/** kmer/impl/CountProcessorHistogram.hpp:104-184 */
template<size_t span> class CountProcessorHistogram : public CountProcessorAbstract<span>
{
public:
    typedef typename Kmer<span>::Type Type;
    CountProcessorHistogram (size_t histoMax = 10000, size_t min_auto_threshold = 3, tools::misc::Histogram* shared = 0)
        : CountProcessorAbstract<span>("histogram"), _histoMax(histoMax), _min_auto_threshold(min_auto_threshold),
          _histogram(shared ? shared : new tools::misc::Histogram (histoMax)), _owns(shared == 0) {}
    ~CountProcessorHistogram () { if (_owns) delete _histogram; }
    ICountProcessor<span>* clone () { return new CountProcessorHistogram (_histoMax, _min_auto_threshold, _histogram); }   /* clones feed the same table (HistogramCache in the reference) */
    void end () { _histogram->compute_threshold ((int)_min_auto_threshold); }
    bool process (size_t, const Type&, const CountVector&, CountNumber sum) { _histogram->inc ((uint32_t)sum); return true; }
    tools::misc::Histogram* getHistogram () { return _histogram; }
private:
    size_t _histoMax, _min_auto_threshold; tools::misc::Histogram* _histogram; bool _owns;
};
/** kmer/impl/CountProcessorSolidity.hpp:163-189 (single bank => kind SUM, ConfigurationAlgorithm.cpp:261-264) */
template<size_t span> class CountProcessorSolidityInfo : public CountProcessorAbstract<span>
{
public:
    typedef typename Kmer<span>::Type Type;
    CountProcessorSolidityInfo (tools::misc::CountRange range = tools::misc::CountRange (2, 0x7fffffff))
        : CountProcessorAbstract<span>("solidity"), _range(range), _total(0), _ok(0) {}
    void begin (const Configuration& c) { if (!c._abundance.empty ()) _range = c._abundance[0]; }
    ICountProcessor<span>* clone () { return new CountProcessorSolidityInfo (_range); }
    void finishClones (std::vector<ICountProcessor<span>*>& clones)
    { for (size_t i = 0; i < clones.size(); i++) if (CountProcessorSolidityInfo* c = dynamic_cast<CountProcessorSolidityInfo*> (clones[i])) { _total += c->_total; _ok += c->_ok; } }
    bool process (size_t, const Type&, const CountVector&, CountNumber sum) { _total++; bool ok = _range.includes (sum); if (ok) _ok++; return ok; }
    uint64_t getNbDistinct () const { return _total; } uint64_t getNbSolid () const { return _ok; } uint64_t getNbWeak () const { return _total - _ok; }
private:
    tools::misc::CountRange _range; uint64_t _total, _ok;
};
/** kmer/impl/CountProcessorDump.hpp:85-152: appends Count{value, abundance} to partition `partId + passId*nb_partitions`
 *  (in-memory partitions here; an HDF5 sink is outside the path) */
template<size_t span> class CountProcessorDump : public CountProcessorAbstract<span>
{
public:
    typedef typename Kmer<span>::Type Type; typedef typename Kmer<span>::Count Count;
    typedef std::vector<std::vector<Count> > Partitions;
    CountProcessorDump (Partitions* shared = 0) : CountProcessorAbstract<span>("dump"), _parts(shared ? shared : new Partitions ()), _owns(shared == 0), _nbPartsPerPass(1), _cur(0) {}
    ~CountProcessorDump () { if (_owns) delete _parts; }
    void begin (const Configuration& c) { _nbPartsPerPass = c._nb_partitions; _parts->assign ((size_t)c._nb_partitions * c._nb_passes, std::vector<Count> ()); }
    ICountProcessor<span>* clone () { CountProcessorDump* c = new CountProcessorDump (_parts); c->_nbPartsPerPass = _nbPartsPerPass; return c; }
    void beginPart (size_t passId, size_t partId, size_t, const char*) { _cur = partId + passId * _nbPartsPerPass; }
    bool process (size_t, const Type& kmer, const CountVector&, CountNumber sum) { (*_parts)[_cur].push_back (Count (kmer, sum)); return true; }
    Partitions& getSolidCounts () { return *_parts; }
    /** fast path used by SortingCountAlgorithm when the chain is the default one: bulk insert of a device-sorted partition */
    void bulk (size_t key, const uint64_t* lo, const uint64_t* hi, const int32_t* counts, uint64_t n)
    { std::vector<Count>& v = (*_parts)[key]; v.reserve (v.size () + n); for (uint64_t i = 0; i < n; i++) v.push_back (Count (Type::make (lo[i], hi ? hi[i] : 0), counts[i])); }
private:
    Partitions* _parts; bool _owns; size_t _nbPartsPerPass, _cur;
};

/********************************************************************************/
/** kmer/impl/SortingCountAlgorithm.hpp:65-263 */
template<size_t span = 32> class SortingCountAlgorithm
{
public:
    typedef typename Kmer<span>::Type Type; typedef typename Kmer<span>::Count Count;
    typedef ICountProcessor<span> CountProcessor;

    /** :102-109 */
    SortingCountAlgorithm (bank::IBank* bank, const Configuration& config, Repartitor* repartitor,
                           std::vector<CountProcessor*> processors = std::vector<CountProcessor*> (), int device = 0)
        : _bank(0), _config(config), _repartitor(0), _device(device), _defaultDump(0), _defaultHisto(0), _defaultSolid(0), _processor(0)
    {
        system::setAttr (_bank, bank); system::setAttr (_repartitor, repartitor);
        for (size_t i = 0; i < processors.size(); i++) addProcessor (processors[i]);
    }
    ~SortingCountAlgorithm ()
    {
        system::setAttr (_bank, (bank::IBank*)0); system::setAttr (_repartitor, (Repartitor*)0);
        for (size_t i = 0; i < _processors.size(); i++) _processors[i]->forget ();
        if (_processor) _processor->forget ();
    }
    /** :171 */
    void addProcessor (CountProcessor* p) { p->use (); _processors.push_back (p); }
    /** :121-150 getDefaultProcessorVector: histogram -> solidity -> dump */
    static std::vector<CountProcessor*> getDefaultProcessorVector (const Configuration& config)
    {
        std::vector<CountProcessor*> v;
        v.push_back (new CountProcessorChain<span> (new CountProcessorHistogram<span> (config._histogramMax),
                                                    new CountProcessorSolidityInfo<span> (config._abundance.empty () ? tools::misc::CountRange (2, 0x7fffffff) : config._abundance[0]),
                                                    new CountProcessorDump<span> ()));
        return v;
    }
    const Configuration& getConfig () const { return _config; }
    Repartitor* getRepartitor () { return _repartitor; }
    tools::misc::Properties& getInfo () { return _info; }
    /** :175 / :179 -- solid k-mers of the default chain, one vector per key pass*nb_partitions+part, ascending */
    typename CountProcessorDump<span>::Partitions* getSolidCounts () { return _defaultDump ? &_defaultDump->getSolidCounts () : 0; }
    tools::misc::Histogram* getHistogram () { return _defaultHisto ? _defaultHisto->getHistogram () : 0; }

    /** :156 -- execute(): SortingCountAlgorithm.cpp:636-781 */
    void execute ()
    {
        if (_config._kmerSize >= span) throw system::Exception ("Type '%s' has too low precision (%d bits) for the required %d kmer size", Type::getName (), (int)Type::getSize (), (int)_config._kmerSize);
        if (_processors.empty ()) { std::vector<CountProcessor*> v = getDefaultProcessorVector (_config); for (size_t i = 0; i < v.size (); i++) addProcessor (v[i]); }
        if (_processor) _processor->forget ();
        _processor = _processors.size () == 1 ? _processors[0] : new CountProcessorChain<span> (_processors);
        _processor->use ();
        // the default chain (exactly histogram -> solidity -> dump) lets the device do histogram + solidity and return solid k-mers only
        _defaultHisto = 0; _defaultSolid = 0; _defaultDump = 0;
        bool fast = false;
        if (CountProcessorChain<span>* chain = dynamic_cast<CountProcessorChain<span>*> (_processor))
            if (chain->items ().size () == 3)
            {
                _defaultHisto = dynamic_cast<CountProcessorHistogram<span>*> (chain->items ()[0]);
                _defaultSolid = dynamic_cast<CountProcessorSolidityInfo<span>*> (chain->items ()[1]);
                _defaultDump  = dynamic_cast<CountProcessorDump<span>*> (chain->items ()[2]);
                fast = _defaultHisto && _defaultSolid && _defaultDump;
            }
        if (!fast) { _defaultHisto = _processor->template get<CountProcessorHistogram<span> > (); _defaultDump = _processor->template get<CountProcessorDump<span> > (); }

        // ---- bank -> one ASCII blob + offsets (the device packs it: gatb_gpu_pack_ascii) ----
        Gather g; _bank->iterate (&Gather::cb, &g);
        g.offsets.push_back (g.blob.size ());
        gatb_gpu_ctx* ctx = gatb_gpu_create (_device);
        if (!ctx) throw system::Exception ("%s", gatb_gpu_last_error (0));
        struct Guard { gatb_gpu_ctx* c; ~Guard () { gatb_gpu_destroy (c); } } guard = { ctx };
        std::vector<uint8_t> packed ((g.blob.size () + 3) / 4 + 64, 0);
        std::vector<uint32_t> nmask ((g.blob.size () + 31) / 32 + 4, 0);
        uint64_t nbad = 0;
        if (gatb_gpu_pack_ascii (ctx, g.blob.data (), g.blob.size (), packed.data (), nmask.data (), &nbad)) throw system::Exception ("%s", gatb_gpu_last_error (ctx));

        gatb_gpu_params p; memset (&p, 0, sizeof(p));
        p.kmer_size = (int32_t)_config._kmerSize; p.minimizer_size = (int32_t)_config._minim_size;
        p.nb_partitions = (int32_t)_config._nb_partitions; p.nb_passes = (int32_t)_config._nb_passes;
        const tools::misc::CountRange range0 = _config._abundance.empty () ? tools::misc::CountRange (2, 0x7fffffff) : _config._abundance[0];
        p.abundance_min = (int32_t)range0.getBegin (); p.abundance_max = (int32_t)std::min<int64_t> (range0.getEnd (), 0x7fffffff);
        p.histo_max = (int32_t)_config._histogramMax; p.minimizer_type = (int32_t)_config._minimizerType; p.emit_all = fast ? 0 : 1;
        const uint16_t* table = (_repartitor && !_repartitor->getTable ().empty ()) ? _repartitor->getTable ().data () : 0;
        gatb_gpu_result r;
        if (gatb_gpu_count (ctx, &p, table, 0, packed.data (), g.offsets.data (), g.offsets.size () - 1, nbad ? nmask.data () : 0, &r))
            throw system::Exception ("%s", gatb_gpu_last_error (ctx));

        // ---- deliver through the ICountProcessor protocol (ICountProcessor.hpp:91-183) ----
        _processor->begin (_config);
        CountVector cv (1);
        for (uint32_t pass = 0; pass < _config._nb_passes; pass++)
        {
            _processor->beginPass (pass);
            std::vector<CountProcessor*> clones;
            for (uint32_t part = 0; part < _config._nb_partitions; part++)
            {
                const uint64_t key = (uint64_t)pass * _config._nb_partitions + part, a = r.part_offsets[key], b = r.part_offsets[key + 1];
                CountProcessor* clone = _processor->clone (); clone->use (); clones.push_back (clone);
                clone->beginPart (pass, part, 0, "device");
                if (fast)
                {   // histogram and solidity were computed on the device; bulk-insert the solid k-mers
                    CountProcessorChain<span>* cc = static_cast<CountProcessorChain<span>*> (clone);
                    static_cast<CountProcessorDump<span>*> (cc->items ()[2])->bulk (key, r.kmers_lo + a, r.kmers_hi ? r.kmers_hi + a : 0, r.counts + a, b - a);
                }
                else
                    for (uint64_t i = a; i < b; i++) { cv[0] = r.counts[i]; clone->process (part, Type::make (r.kmers_lo[i], r.kmers_hi ? r.kmers_hi[i] : 0), cv, r.counts[i]); }
                clone->endPart (pass, part);
            }
            _processor->finishClones (clones);
            for (size_t i = 0; i < clones.size (); i++) clones[i]->forget ();
            _processor->endPass (pass);
        }
        if (fast) { _defaultHisto->getHistogram ()->set (r.histogram); }
        _processor->end ();
        /** SortingCountAlgorithm.cpp:728-780 */
        _info.add ("kmers_nb_valid", r.stats[GATB_STAT_KMERS_VALID]); _info.add ("kmers_nb_invalid", r.stats[GATB_STAT_KMERS_INVALID]);
        _info.add ("kmers_nb_distinct", r.stats[GATB_STAT_DISTINCT]); _info.add ("kmers_nb_solid", r.stats[GATB_STAT_SOLID]);
        _info.add ("kmers_nb_weak", r.stats[GATB_STAT_DISTINCT] - r.stats[GATB_STAT_SOLID]);
        _info.add ("sequences_number", r.stats[GATB_STAT_SEQUENCES]); _info.add ("sequences_size", r.stats[GATB_STAT_NUCLEOTIDES]);
        _info.add ("nb_partitions", _config._nb_partitions); _info.add ("nb_passes", _config._nb_passes);
        gatb_gpu_result_free (ctx, &r);
    }
private:
    struct Gather
    {
        std::string blob; std::vector<uint64_t> offsets;
        static void cb (const bank::Sequence& s, void* self) { Gather* g = (Gather*)self; g->offsets.push_back (g->blob.size ()); g->blob += s.data; }
    };
    bank::IBank* _bank; Configuration _config; Repartitor* _repartitor; int _device;
    std::vector<CountProcessor*> _processors;
    CountProcessorDump<span>* _defaultDump; CountProcessorHistogram<span>* _defaultHisto; CountProcessorSolidityInfo<span>* _defaultSolid;
    CountProcessor* _processor; tools::misc::Properties _info;
};

/********************************************************************************/
/** kmer/impl/BloomBuilder.hpp:102-131 + BloomAlgorithm.cpp:155-203: Bloom filter of the solid k-mers, built on the device */
template<size_t span = 32> class BloomBuilder
{
public:
    typedef typename Kmer<span>::Type Type; typedef typename Kmer<span>::Count Count;
    /** kind: "basic", "cache", "neighbor" (tools/misc/api/Enums.hpp BloomKind) */
    BloomBuilder (uint64_t bloomSize, size_t nbHash, size_t kmerSize, const std::string& kind = "neighbor", int device = 0)
        : _bloomSize(bloomSize), _nbHash(nbHash), _kmerSize(kmerSize), _device(device)
    {
        if (kind == "basic") _kind = GATB_BLOOM_BASIC; else if (kind == "cache") _kind = GATB_BLOOM_CACHE; else if (kind == "neighbor") _kind = GATB_BLOOM_NEIGHBOR;
        else throw system::Exception ("bad Bloom kind '%s' in createBloom", kind.c_str ());
    }
    /** BloomAlgorithm.cpp:158-166 */
    static void sizeFor (size_t kmerSize, uint64_t nbSolid, uint64_t& bloomSize, size_t& nbHash)
    { int32_t h; if (gatb_gpu_bloom_params ((int)kmerSize, nbSolid, &bloomSize, &h)) throw system::Exception ("bad kmer size %d", (int)kmerSize); nbHash = h; }
    /** returns the byte array (1 + tai/8 bytes) exactly as IBloom::getArray(); bitSize = IBloom::getBitSize() */
    std::vector<uint8_t> build (const std::vector<Count>& solid, uint64_t* bitSize = 0)
    {
        std::vector<uint64_t> lo (solid.size ()), hi (solid.size ());
        for (size_t i = 0; i < solid.size (); i++) { lo[i] = solid[i].value.lo (); hi[i] = solid[i].value.hi (); }
        uint64_t nbytes = 0, bits = 0; gatb_gpu_bloom_layout (_kind, _bloomSize, &nbytes, &bits);
        if (bitSize) *bitSize = bits;
        std::vector<uint8_t> out (nbytes, 0);
        gatb_gpu_ctx* ctx = gatb_gpu_create (_device);
        if (!ctx) throw system::Exception ("%s", gatb_gpu_last_error (0));
        int rc = gatb_gpu_bloom (ctx, _kind, _bloomSize, (int)_nbHash, (int)_kmerSize, lo.data (), span > 32 ? hi.data () : 0, lo.size (), out.data ());
        std::string err = rc ? gatb_gpu_last_error (ctx) : "";
        gatb_gpu_destroy (ctx);
        if (rc) throw system::Exception ("%s", err.c_str ());
        return out;
    }
private:
    uint64_t _bloomSize; size_t _nbHash, _kmerSize; int _kind, _device;
};

} } } } // impl, kmer, core, gatb
#endif
