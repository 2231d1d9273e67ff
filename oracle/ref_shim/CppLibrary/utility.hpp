// Stand-in for the un-vendored third-party dependency "Cpp-Library" (namespace CL) that
// the reference includes as <CppLibrary/utility.hpp> (reference README.md:59-61,
// CMakeLists.txt:15-17; no version is pinned anywhere in the reference).
//
// TEST INFRASTRUCTURE ONLY: used to compile the reference's own sources into
// oracle/_ref/ so the oracle can be pinned against the real implementation.
// It contributes no arithmetic to the hot path - only containers, string
// splitting and file helpers (call sites: source/IC/InvDisp.cpp:859,
// source/IC/IntCoordSet.cpp:16,61, source/polynomial/SAP.cpp:47,
// source/polynomial/SAPSet.cpp:92,97).
#ifndef CL_SHIM_UTILITY_HPP
#define CL_SHIM_UTILITY_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <forward_list>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace CL { namespace utility {

class file_error : public std::runtime_error {
public:
    explicit file_error(const std::string & file)
    : std::runtime_error("CL::utility::file_error: cannot open " + file) {}
};

class not_ready : public std::logic_error {
public:
    explicit not_ready(const std::string & what)
    : std::logic_error("CL::utility::not_ready: " + what) {}
};

// dense rows x cols container addressed as m[i][j]
template <typename T> class matrix {
    std::vector<std::vector<T>> rows_;
public:
    matrix() {}
    explicit matrix(const size_t & n) : rows_(n, std::vector<T>(n)) {}
    matrix(const size_t & n, const size_t & m) : rows_(n, std::vector<T>(m)) {}
    size_t size(const size_t & dim = 0) const {
        if (dim == 0) return rows_.size();
        return rows_.empty() ? 0 : rows_[0].size();
    }
    std::vector<T> & operator[](const size_t & i) {return rows_[i];}
    const std::vector<T> & operator[](const size_t & i) const {return rows_[i];}
    void resize(const size_t & n) {
        rows_.resize(n);
        for (auto & row : rows_) row.resize(n);
    }
    void resize(const size_t & n, const size_t & m) {
        rows_.resize(n);
        for (auto & row : rows_) row.resize(m);
    }
    typename std::vector<std::vector<T>>::iterator begin() {return rows_.begin();}
    typename std::vector<std::vector<T>>::iterator end() {return rows_.end();}
    typename std::vector<std::vector<T>>::const_iterator begin() const {return rows_.begin();}
    typename std::vector<std::vector<T>>::const_iterator end() const {return rows_.end();}
};

// whitespace split
inline std::vector<std::string> split(const std::string & text) {
    std::istringstream iss(text);
    return std::vector<std::string>(std::istream_iterator<std::string>{iss},
                                    std::istream_iterator<std::string>());
}
// single-character split, empty pieces dropped
inline std::vector<std::string> split(const std::string & text, const char & sep) {
    std::vector<std::string> out;
    std::string piece;
    std::istringstream iss(text);
    while (std::getline(iss, piece, sep)) if (! piece.empty()) out.push_back(piece);
    return out;
}

inline std::vector<double> read_vector(std::ifstream & ifs) {
    std::vector<double> data;
    double x;
    while (ifs >> x) data.push_back(x);
    return data;
}
inline std::vector<double> read_vector(const std::string & file) {
    std::ifstream ifs(file);
    if (! ifs.good()) throw file_error(file);
    return read_vector(ifs);
}

} // namespace utility
} // namespace CL

#endif
