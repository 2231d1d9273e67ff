// tchem::utility - drop-in for include/tchem/utility.hpp of Torch-Chemistry (source/utility.cpp): small host
// helpers the reference's programs use around the internal-coordinate path (reading a vector from a text file,
// tensor <-> std::vector conversions, parameter counting). Set-up time host code: no device work.
#ifndef tchem_utility_hpp
#define tchem_utility_hpp

#include <fstream>

#include <torch/torch.h>

#include <CppLibrary/utility.hpp>

namespace tchem { namespace utility {

std::vector<double> tensor2vector(const at::Tensor & tensor);
CL::utility::matrix<double> tensor2matrix(const at::Tensor & tensor);

// whitespace-separated numbers of a text file as a float64 vector
at::Tensor read_vector(std::ifstream & ifs);
at::Tensor read_vector(const std::string & file);

size_t NParameters(const std::vector<at::Tensor> & parameters);
double NetGradNorm(const std::vector<at::Tensor> & parameters);

} }

#endif
