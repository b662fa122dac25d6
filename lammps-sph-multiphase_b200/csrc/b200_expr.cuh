// b200_expr.cuh -- atom-style / equal-style variable formulas on the device.
//
// fix addforce and fix setmeso take `v_name` arguments whose value the reference evaluates for every atom on every step
// (FixAddForce::post_force -> Variable::compute_atom, fix_addforce.cpp:288-320; fix_setmeso.cpp:238-262).  The shipped decks use
// them for body forces (`mass*${gx}`, poiseuille.lmp:57 `mass*${gx}*((y<${Ly}/2.0)-(y>${Ly}/2.0))`) and for initial profiles.
// The formula text is compiled on the host into a postfix program with the reference's own operator table and pop rule
// (Variable::evaluate, variable.cpp:99-107,1641: every binary operator, `^` included, associates to the left; unary minus and
// `!` bind tightest), and the device walks that program per atom with the same IEEE operations the reference's tree walk performs
// (Variable::eval_tree, variable.cpp:2090-2500).  Supported: numbers, PI, the atom vectors id mass type x y z vx vy vz fx fy fz, the
// thermo keywords step and dt, + - * / % ^ == != < <= > >= && || ! and the one- and two-argument math functions below.
// Anything else (group / region functions, computes, fixes, random()) is refused with a message when the fix is registered.
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#define EXPR_MAXOPS 64
#define EXPR_STACK 16
enum {
  EX_CONST = 0, EX_ID, EX_MASS, EX_TYPE, EX_X, EX_Y, EX_Z, EX_VX, EX_VY, EX_VZ, EX_FX, EX_FY, EX_FZ, EX_STEP, EX_DT, EX_TIME,
  EX_ADD, EX_SUB, EX_MUL, EX_DIV, EX_MOD, EX_POW, EX_NEG, EX_NOT, EX_EQ, EX_NE, EX_LT, EX_LE, EX_GT, EX_GE, EX_AND, EX_OR,
  EX_SQRT, EX_EXP, EX_LN, EX_LOG, EX_ABS, EX_SIN, EX_COS, EX_TAN, EX_ASIN, EX_ACOS, EX_ATAN, EX_ATAN2, EX_CEIL, EX_FLOOR, EX_ROUND
};
struct ExprProg { int n; unsigned char op[EXPR_MAXOPS]; double val[EXPR_MAXOPS]; };
struct ExprIn { double x, y, z, vx, vy, vz, fx, fy, fz, mass, step, dt, time; int type, id; };

static __host__ __device__ inline double expr_eval(const ExprProg &P, const ExprIn &in)
{
  double st[EXPR_STACK]; int sp = 0;
  for (int k = 0; k < P.n; k++) {
    const int op = P.op[k];
    if (op < EX_ADD) {
      double v;
      switch (op) {
      case EX_CONST: v = P.val[k]; break;
      case EX_ID: v = in.id; break;
      case EX_MASS: v = in.mass; break;
      case EX_TYPE: v = in.type; break;
      case EX_X: v = in.x; break;
      case EX_Y: v = in.y; break;
      case EX_Z: v = in.z; break;
      case EX_VX: v = in.vx; break;
      case EX_VY: v = in.vy; break;
      case EX_VZ: v = in.vz; break;
      case EX_FX: v = in.fx; break;
      case EX_FY: v = in.fy; break;
      case EX_FZ: v = in.fz; break;
      case EX_STEP: v = in.step; break;
      case EX_DT: v = in.dt; break;
      default: v = in.time; break;
      }
      st[sp++] = v;
      continue;
    }
    const bool binary = (op <= EX_OR && op != EX_NEG && op != EX_NOT) || op == EX_ATAN2;
    double b = st[--sp], a = 0.0;
    if (binary) a = st[--sp];
    double r;
    switch (op) {
    case EX_ADD: r = a + b; break;
    case EX_SUB: r = a - b; break;
    case EX_MUL: r = a * b; break;
    case EX_DIV: r = a / b; break;
    case EX_MOD: r = fmod(a, b); break;
    case EX_POW: r = pow(a, b); break;
    case EX_NEG: r = -b; break;
    case EX_NOT: r = b == 0.0 ? 1.0 : 0.0; break;
    case EX_EQ: r = a == b ? 1.0 : 0.0; break;
    case EX_NE: r = a != b ? 1.0 : 0.0; break;
    case EX_LT: r = a < b ? 1.0 : 0.0; break;
    case EX_LE: r = a <= b ? 1.0 : 0.0; break;
    case EX_GT: r = a > b ? 1.0 : 0.0; break;
    case EX_GE: r = a >= b ? 1.0 : 0.0; break;
    case EX_AND: r = (a != 0.0 && b != 0.0) ? 1.0 : 0.0; break;
    case EX_OR: r = (a != 0.0 || b != 0.0) ? 1.0 : 0.0; break;
    case EX_SQRT: r = sqrt(b); break;
    case EX_EXP: r = exp(b); break;
    case EX_LN: r = log(b); break;
    case EX_LOG: r = log10(b); break;
    case EX_ABS: r = fabs(b); break;
    case EX_SIN: r = sin(b); break;
    case EX_COS: r = cos(b); break;
    case EX_TAN: r = tan(b); break;
    case EX_ASIN: r = asin(b); break;
    case EX_ACOS: r = acos(b); break;
    case EX_ATAN: r = atan(b); break;
    case EX_ATAN2: r = atan2(a, b); break;
    case EX_CEIL: r = ceil(b); break;
    case EX_FLOOR: r = floor(b); break;
    default: r = (b < 0.0) ? ceil(b - 0.5) : floor(b + 0.5); break;      // MYROUND, variable.cpp:55
    }
    st[sp++] = r;
  }
  return sp ? st[sp - 1] : 0.0;
}

// ---- host: formula text -> postfix program (the operator-precedence loop of Variable::evaluate) ----
struct ExprCompiler {
  const char *s; size_t i = 0; ExprProg &P; std::string err; int depth = 0, maxdepth = 0;
  ExprCompiler(const char *str, ExprProg &p) : s(str), P(p) { P.n = 0; }
  bool emit(int op, double v = 0.0)
  {
    if (P.n >= EXPR_MAXOPS) { err = "variable formula too long for the device evaluator"; return false; }
    P.op[P.n] = (unsigned char)op; P.val[P.n] = v; P.n++;
    if (op < EX_ADD) { if (++depth > maxdepth) maxdepth = depth; }
    else if ((op <= EX_OR && op != EX_NEG && op != EX_NOT) || op == EX_ATAN2) depth--;
    if (maxdepth > EXPR_STACK) { err = "variable formula nests too deep for the device evaluator"; return false; }
    return true;
  }
  static int prec(int op)
  {
    switch (op) {
    case EX_OR: return 1; case EX_AND: return 2; case EX_EQ: case EX_NE: return 3;
    case EX_LT: case EX_LE: case EX_GT: case EX_GE: return 4; case EX_ADD: case EX_SUB: return 5;
    case EX_MUL: case EX_DIV: case EX_MOD: return 6; case EX_POW: return 7; default: return 8;      // unary minus, not
    }
  }
  // formula up to the closing ')' or ',' of the enclosing call, or the end of the string
  bool formula(const char *stops)
  {
    std::vector<int> ops;
    bool expect_arg = true;
    for (;;) {
      while (s[i] == ' ' || s[i] == '\t') i++;
      const char c = s[i];
      if (expect_arg) {
        if (c == '-') { ops.push_back(EX_NEG); i++; continue; }
        if (c == '!' && s[i + 1] != '=') { ops.push_back(EX_NOT); i++; continue; }
        if (c == '(') {
          i++;
          if (!formula(")")) return false;
          if (s[i] != ')') { err = "Invalid syntax in variable formula"; return false; }
          i++; expect_arg = false; continue;
        }
        if ((c >= '0' && c <= '9') || c == '.') {
          size_t j = i;
          while ((s[j] >= '0' && s[j] <= '9') || s[j] == '.' || s[j] == 'e' || s[j] == 'E' || ((s[j] == '-' || s[j] == '+') && j > i && (s[j - 1] == 'e' || s[j - 1] == 'E'))) j++;
          if (!emit(EX_CONST, atof(std::string(s + i, j - i).c_str()))) return false;
          i = j; expect_arg = false; continue;
        }
        if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_') {
          size_t j = i;
          while ((s[j] >= 'a' && s[j] <= 'z') || (s[j] >= 'A' && s[j] <= 'Z') || (s[j] >= '0' && s[j] <= '9') || s[j] == '_') j++;
          const std::string w(s + i, j - i);
          i = j;
          if (s[i] == '(') {
            static const struct { const char *n; int op, nargs; } F[] = {
              {"sqrt", EX_SQRT, 1}, {"exp", EX_EXP, 1}, {"ln", EX_LN, 1}, {"log", EX_LOG, 1}, {"abs", EX_ABS, 1}, {"sin", EX_SIN, 1}, {"cos", EX_COS, 1},
              {"tan", EX_TAN, 1}, {"asin", EX_ASIN, 1}, {"acos", EX_ACOS, 1}, {"atan", EX_ATAN, 1}, {"atan2", EX_ATAN2, 2}, {"ceil", EX_CEIL, 1},
              {"floor", EX_FLOOR, 1}, {"round", EX_ROUND, 1}};
            int f = -1;
            for (int k = 0; k < (int)(sizeof F / sizeof F[0]); k++) if (w == F[k].n) f = k;
            if (f < 0) { err = "variable function " + w + "() is not available in the /b200 fixes"; return false; }
            i++;
            for (int a = 0; a < F[f].nargs; a++) {
              if (!formula(a + 1 < F[f].nargs ? "," : ")")) return false;
              if (s[i] != (a + 1 < F[f].nargs ? ',' : ')')) { err = "Invalid math function in variable formula"; return false; }
              i++;
            }
            if (!emit(F[f].op)) return false;
            expect_arg = false; continue;
          }
          static const struct { const char *n; int op; } V[] = {
            {"id", EX_ID}, {"mass", EX_MASS}, {"type", EX_TYPE}, {"x", EX_X}, {"y", EX_Y}, {"z", EX_Z}, {"vx", EX_VX}, {"vy", EX_VY}, {"vz", EX_VZ},
            {"fx", EX_FX}, {"fy", EX_FY}, {"fz", EX_FZ}, {"step", EX_STEP}, {"dt", EX_DT}};
          int v = -1;
          for (int k = 0; k < (int)(sizeof V / sizeof V[0]); k++) if (w == V[k].n) v = V[k].op;
          if (w == "PI") { if (!emit(EX_CONST, 3.14159265358979323846)) return false; }
          else if (v >= 0) { if (!emit(v)) return false; }
          else { err = "variable keyword '" + w + "' is not available in the /b200 fixes"; return false; }
          expect_arg = false; continue;
        }
        err = "Invalid syntax in variable formula"; return false;
      }
      // an operator, or the end of this (sub)formula
      int op = -1; size_t len = 1;
      if (c == 0 || strchr(stops, c)) op = -2;
      else if (c == '+') op = EX_ADD; else if (c == '-') op = EX_SUB; else if (c == '*') op = EX_MUL; else if (c == '/') op = EX_DIV;
      else if (c == '%') op = EX_MOD; else if (c == '^') op = EX_POW;
      else if (c == '=' && s[i + 1] == '=') { op = EX_EQ; len = 2; } else if (c == '!' && s[i + 1] == '=') { op = EX_NE; len = 2; }
      else if (c == '<' && s[i + 1] == '=') { op = EX_LE; len = 2; } else if (c == '<') op = EX_LT;
      else if (c == '>' && s[i + 1] == '=') { op = EX_GE; len = 2; } else if (c == '>') op = EX_GT;
      else if (c == '&' && s[i + 1] == '&') { op = EX_AND; len = 2; } else if (c == '|' && s[i + 1] == '|') { op = EX_OR; len = 2; }
      if (op == -1) { err = "Invalid syntax in variable formula"; return false; }
      const int p = op == -2 ? 0 : prec(op);
      while (!ops.empty() && prec(ops.back()) >= p) { if (!emit(ops.back())) return false; ops.pop_back(); }      // variable.cpp:1641
      if (op == -2) return true;
      ops.push_back(op); i += len; expect_arg = true;
    }
  }
};
// returns an empty string on success
static inline std::string expr_compile(const char *text, ExprProg &P)
{
  ExprCompiler C(text, P);
  if (!C.formula("")) return C.err.empty() ? std::string("Invalid syntax in variable formula") : C.err;
  if (C.s[C.i] != 0 || C.depth != 1) return "Invalid syntax in variable formula";
  return "";
}
