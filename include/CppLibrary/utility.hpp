// Minimal stand-in for the parts of the third-party "Cpp-Library" (namespace CL) that appear
// in tchem's public interface on the internal-coordinate path: the file_error exception thrown
// by the definition-file readers and the matrix<T> container. The real library is an external,
// un-vendored dependency of Torch-Chemistry; when it is installed, put it first on the include
// path and this header is never seen.
#ifndef CL_utility_hpp
#define CL_utility_hpp

#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace CL { namespace utility {

// thrown by the observables of an analysis whose kernel() has not run (reference source/chem/normal_mode.cpp:79-88)
class not_ready : public std::runtime_error {
    public:
    explicit not_ready(const std::string & what) : std::runtime_error("CL::utility::not_ready: " + what) {}
};
class file_error : public std::runtime_error {
public:
    explicit file_error(const std::string & file) : std::runtime_error("CL::utility::file_error: " + file) {}
};

template <typename T> class matrix {
    std::vector<std::vector<T>> data_;
public:
    matrix() {}
    explicit matrix(const size_t & n) : data_(n, std::vector<T>(n)) {}
    matrix(const size_t & rows, const size_t & cols) : data_(rows, std::vector<T>(cols)) {}
    size_t size(const size_t & dim = 0) const { return dim == 0 ? data_.size() : (data_.empty() ? 0 : data_[0].size()); }
    std::vector<T> & operator[](const size_t & i) { return data_[i]; }
    const std::vector<T> & operator[](const size_t & i) const { return data_[i]; }
};

inline std::vector<std::string> split(const std::string & text) {
    std::istringstream iss(text);
    std::vector<std::string> out;
    std::string tok;
    while (iss >> tok) out.push_back(tok);
    return out;
}

} }

#endif
