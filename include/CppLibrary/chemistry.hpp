// Stand-in for the CL::chem geometry readers of the un-vendored Cpp-Library that tchem host code uses
// (test/intcoord/main.cpp:14,55,69,123,127): .xyz files and Columbus geom files. Containers / file parsing only.
#ifndef CL_chemistry_hpp
#define CL_chemistry_hpp

#include <fstream>
#include <string>
#include <vector>

#include <CppLibrary/utility.hpp>

namespace CL { namespace chem {

// standard .xyz: natoms / comment / symbol x y z ; Angstrom, converted to bohr when atomic_unit
template <typename T> class xyz {
protected:
    std::vector<std::string> symbols_;
    std::vector<T> coords_;
public:
    xyz() {}
    xyz(const std::string & file, const bool & atomic_unit = false) {
        std::ifstream ifs(file);
        if (! ifs.good()) throw CL::utility::file_error(file);
        std::string line;
        std::getline(ifs, line);
        size_t natoms = std::stoul(line);
        std::getline(ifs, line);
        for (size_t i = 0; i < natoms; i++) {
            std::string s; T x, y, z;
            ifs >> s >> x >> y >> z;
            symbols_.push_back(s);
            coords_.push_back(x); coords_.push_back(y); coords_.push_back(z);
        }
        if (atomic_unit) for (T & c : coords_) c *= 1.8897261339212517;
    }
    size_t NAtoms() const {return symbols_.size();}
    const std::vector<std::string> & symbols() const {return symbols_;}
    const std::vector<T> & coords() const {return coords_;}
};

// Columbus geom: symbol Z x y z mass, already in bohr
template <typename T> class xyz_mass : public xyz<T> {
    std::vector<T> masses_;
public:
    xyz_mass() {}
    xyz_mass(const std::string & file, const bool & atomic_unit = false) {
        (void)atomic_unit;
        std::ifstream ifs(file);
        if (! ifs.good()) throw CL::utility::file_error(file);
        std::string s; T Z, x, y, z, m;
        while (ifs >> s >> Z >> x >> y >> z >> m) {
            this->symbols_.push_back(s);
            this->coords_.push_back(x); this->coords_.push_back(y); this->coords_.push_back(z);
            masses_.push_back(m);
        }
    }
    const std::vector<T> & masses() const {return masses_;}
};

} }
#endif
