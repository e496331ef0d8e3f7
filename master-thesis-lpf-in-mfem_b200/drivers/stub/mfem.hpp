// STUB of the MFEM declarations used by include/lpf_mfem_adapter.hpp -- compile check only.
// MFEM itself is not available in this image.  Only signatures are declared, mirroring
// upstream MFEM 4.x ([MFEM] general/array.hpp, linalg/vector.hpp, linalg/operator.hpp, linalg/ode.hpp,
// fem/bilininteg.hpp, fem/fespace.hpp, fem/restriction.hpp, mesh/mesh.hpp); nothing here computes.
#pragma once
#include <cstdio>
#include <cstdlib>

namespace mfem {

using real_t = double;
inline void mfem_error(const char *msg) { std::fprintf(stderr, "MFEM abort: %s\n", msg); std::abort(); }

template <class T>
class Array {
public:
    Array() = default;
    Array(T *d, int n) : data_(d), size_(n) {}
    int Size() const { return size_; }
    T *GetData() const { return data_; }
    const T *HostRead() const { return data_; }
private:
    T *data_ = nullptr;
    int size_ = 0;
};

class Vector {
public:
    Vector() = default;
    Vector(double *d, int n) : data_(d), size_(n) {}
    int Size() const { return size_; }
    double *GetData() const { return data_; }
    const double *Read() const { return data_; }         // device pointer under Device("cuda")
    double *Write() { return data_; }
    double *ReadWrite() { return data_; }
    const double *HostRead() const { return data_; }
    double *HostWrite() { return data_; }
private:
    double *data_ = nullptr;
    int size_ = 0;
};

class Operator {
public:
    explicit Operator(int s = 0) : height(s), width(s) {}
    virtual ~Operator() = default;
    int Height() const { return height; }
    int Width() const { return width; }
    virtual void Mult(const Vector &x, Vector &y) const = 0;
    virtual void AssembleDiagonal(Vector &) const { mfem_error("not supported"); }
protected:
    int height, width;
};

class Solver : public Operator {
public:
    explicit Solver(int s = 0, bool iter_mode = false) : Operator(s), iterative_mode(iter_mode) {}
    virtual void SetOperator(const Operator &op) = 0;
    bool iterative_mode;
};

class TimeDependentOperator : public Operator {
public:
    explicit TimeDependentOperator(int n = 0, double t_ = 0.0) : Operator(n), t(t_) {}
    virtual double GetTime() const { return t; }
    virtual void SetTime(const double t_) { t = t_; }
protected:
    double t;
};

class ODESolver {
public:
    virtual ~ODESolver() = default;
    virtual void Init(TimeDependentOperator &f_) { f = &f_; }
    virtual void Step(Vector &x, double &t, double &dt) = 0;
protected:
    TimeDependentOperator *f = nullptr;
};

struct Geometry { enum Type { CUBE = 5 }; };
class IntegrationRule {};
class IntegrationRules {
public:
    const IntegrationRule &Get(int, int) { static IntegrationRule ir; return ir; }
};
static IntegrationRules IntRules;

class GeometricFactors {
public:
    enum FactorFlags { COORDINATES = 1, JACOBIANS = 2, DETERMINANTS = 4 };
    Vector J;
};

class IntegrationPoint {
public:
    double x = 0, y = 0, z = 0;
    void Set3(double a, double b, double c) { x = a; y = b; z = c; }
};
class ElementTransformation {
public:
    void Transform(const IntegrationPoint &, Vector &) {}
};
class Mesh {
public:
    const GeometricFactors *GetGeometricFactors(const IntegrationRule &, int) { static GeometricFactors g; return &g; }
    ElementTransformation *GetElementTransformation(int) { static ElementTransformation t; return &t; }
};

class FiniteElement {
public:
    int GetOrder() const { return 1; }
    Geometry::Type GetGeomType() const { return Geometry::CUBE; }
};

enum class ElementDofOrdering { NATIVE, LEXICOGRAPHIC };

class ElementRestriction : public Operator {
public:
    void Mult(const Vector &, Vector &) const override {}
    const Array<int> &GatherMap() const { return gather_; }
private:
    Array<int> gather_;
};

class FiniteElementSpace {
public:
    const FiniteElement *GetFE(int) const { static FiniteElement fe; return &fe; }
    int GetNE() const { return 0; }
    int GetVSize() const { return 0; }
    int GetTrueVSize() const { return 0; }
    Mesh *GetMesh() const { static Mesh m; return &m; }
    const Operator *GetElementRestriction(ElementDofOrdering) const { static ElementRestriction r; return &r; }
};

class BilinearFormIntegrator {
public:
    virtual ~BilinearFormIntegrator() = default;
    virtual void AssemblePA(const FiniteElementSpace &) { mfem_error("not supported"); }
    virtual void AddMultPA(const Vector &, Vector &) const { mfem_error("not supported"); }
    virtual void AssembleDiagonalPA(Vector &) { mfem_error("not supported"); }
};

}  // namespace mfem
