/*
 * ast.c -- AST nodes, SBML Level-1 infix parser and formula printer.
 *
 * Stands in for the libSBML pieces the reference links against
 * (ASTNode.cpp, FormulaParser.c, FormulaFormatter.c).  Behaviour that the
 * reference's unit tests pin (unittest/test_odeModel.c:155-358 equation
 * strings, unittest/test_processAST.c:190-367) is reproduced: operator
 * precedence and grouping rules, "%.15g" reals, function spellings
 * (acos/asin/atan/ceil/log/pow), log10()/sqrt() sugar, unary minus folding
 * into numbers.
 */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/ast.h"

/* ------------------------------------------------------------------ list */
List_t *List_create(void) { return (List_t *)calloc(1, sizeof(List_t)); }
void List_add(List_t *l, void *p)
{
  if (l->size == l->cap) {
    l->cap = l->cap ? 2 * l->cap : 8;
    l->items = (void **)realloc(l->items, l->cap * sizeof(void *));
  }
  l->items[l->size++] = p;
}
void *List_get(const List_t *l, unsigned int i) { return i < l->size ? l->items[i] : NULL; }
void *List_remove(List_t *l, unsigned int i)
{
  void *p;
  if (i >= l->size) return NULL;
  p = l->items[i];
  memmove(l->items + i, l->items + i + 1, (l->size - i - 1) * sizeof(void *));
  l->size--;
  return p;
}
unsigned int List_size(const List_t *l) { return l ? l->size : 0; }
void List_free(List_t *l) { if (l) { free(l->items); free(l); } }

/* ----------------------------------------------------------------- nodes */
static char *dupstr(const char *s)
{
  char *r;
  if (!s) return NULL;
  r = (char *)malloc(strlen(s) + 1);
  strcpy(r, s);
  return r;
}

ASTNode_t *ASTNode_create(void)
{
  ASTNode_t *n = (ASTNode_t *)calloc(1, sizeof(ASTNode_t));
  n->type = AST_UNKNOWN;
  n->denom = 1;
  n->index = -1;
  return n;
}
ASTNode_t *ASTNode_createWithType(ASTNodeType_t t) { ASTNode_t *n = ASTNode_create(); n->type = t; return n; }
ASTNode_t *ASTNode_createIndexName(void) { ASTNode_t *n = ASTNode_create(); n->type = AST_NAME; n->index = 0; return n; }

void ASTNode_free(ASTNode_t *n)
{
  unsigned int i;
  if (!n) return;
  for (i = 0; i < n->nchildren; i++) ASTNode_free(n->children[i]);
  free(n->children);
  free(n->name);
  free(n);
}

ASTNode_t *ASTNode_deepCopy(const ASTNode_t *f)
{
  unsigned int i;
  ASTNode_t *n;
  if (!f) return NULL;
  n = ASTNode_create();
  n->type = f->type; n->ival = f->ival; n->denom = f->denom; n->rval = f->rval;
  n->name = dupstr(f->name); n->index = f->index; n->isData = f->isData;
  for (i = 0; i < f->nchildren; i++) ASTNode_addChild(n, ASTNode_deepCopy(f->children[i]));
  return n;
}

void ASTNode_addChild(ASTNode_t *n, ASTNode_t *c)
{
  if (n->nchildren == n->cap) {
    n->cap = n->cap ? 2 * n->cap : 2;
    n->children = (ASTNode_t **)realloc(n->children, n->cap * sizeof(ASTNode_t *));
  }
  n->children[n->nchildren++] = c;
}
void ASTNode_prependChild(ASTNode_t *n, ASTNode_t *c)
{
  ASTNode_addChild(n, c);
  memmove(n->children + 1, n->children, (n->nchildren - 1) * sizeof(ASTNode_t *));
  n->children[0] = c;
}
ASTNode_t *ASTNode_getChild(const ASTNode_t *n, unsigned int i) { return (n && i < n->nchildren) ? n->children[i] : NULL; }
ASTNode_t *ASTNode_getLeftChild(const ASTNode_t *n) { return ASTNode_getChild(n, 0); }
ASTNode_t *ASTNode_getRightChild(const ASTNode_t *n)
{
  /* libSBML: the last child, but only when there are at least two */
  return (n && n->nchildren > 1) ? n->children[n->nchildren - 1] : NULL;
}
unsigned int ASTNode_getNumChildren(const ASTNode_t *n) { return n ? n->nchildren : 0; }
void ASTNode_swapChildren(ASTNode_t *a, ASTNode_t *b)
{
  ASTNode_t **c = a->children; unsigned int n = a->nchildren, cap = a->cap;
  a->children = b->children; a->nchildren = b->nchildren; a->cap = b->cap;
  b->children = c; b->nchildren = n; b->cap = cap;
}

/* n-ary plus/times -> left-associated binary tree: (((a op b) op c) op d) */
void ASTNode_reduceToBinary(ASTNode_t *n)
{
  unsigned int i;
  if (!n) return;
  for (i = 0; i < n->nchildren; i++) ASTNode_reduceToBinary(n->children[i]);
  if ((n->type == AST_PLUS || n->type == AST_TIMES) && n->nchildren > 2) {
    ASTNode_t *acc = n->children[0];
    for (i = 1; i + 1 < n->nchildren; i++) {
      ASTNode_t *op = ASTNode_createWithType(n->type);
      ASTNode_addChild(op, acc);
      ASTNode_addChild(op, n->children[i]);
      acc = op;
    }
    n->children[0] = acc;
    n->children[1] = n->children[n->nchildren - 1];
    n->nchildren = 2;
  }
}

ASTNodeType_t ASTNode_getType(const ASTNode_t *n) { return n ? n->type : AST_UNKNOWN; }
void ASTNode_setType(ASTNode_t *n, ASTNodeType_t t)
{
  /* as in libSBML: switching between name-ish kinds keeps the name,
     switching to an operator/number drops it */
  int namekind = (t == AST_NAME || t == AST_NAME_TIME || t == AST_FUNCTION || t == AST_FUNCTION_DELAY);
  if (!namekind) { free(n->name); n->name = NULL; n->index = -1; n->isData = 0; }
  n->type = t;
}
void ASTNode_setCharacter(ASTNode_t *n, char c)
{
  switch (c) { case '+': case '-': case '*': case '/': case '^': ASTNode_setType(n, (ASTNodeType_t)c); break;
               default: ASTNode_setType(n, AST_UNKNOWN); }
}
void ASTNode_setName(ASTNode_t *n, const char *name)
{
  char *d = dupstr(name);
  if (ASTNode_isOperator(n) || ASTNode_isNumber(n) || n->type == AST_UNKNOWN) n->type = AST_NAME;
  free(n->name);
  n->name = d;
}
static void drop_children(ASTNode_t *n)
{
  unsigned int i;
  for (i = 0; i < n->nchildren; i++) ASTNode_free(n->children[i]);
  n->nchildren = 0;
}
void ASTNode_setInteger(ASTNode_t *n, long v) { ASTNode_setType(n, AST_INTEGER); n->ival = v; n->rval = 0; drop_children(n); }
void ASTNode_setReal(ASTNode_t *n, double v) { ASTNode_setType(n, AST_REAL); n->rval = v; n->ival = 0; drop_children(n); }
void ASTNode_setRealWithExponent(ASTNode_t *n, double m, long e)
{ ASTNode_setType(n, AST_REAL_E); n->rval = m * pow(10.0, (double)e); n->ival = e; drop_children(n); }
void ASTNode_setRational(ASTNode_t *n, long a, long b) { ASTNode_setType(n, AST_RATIONAL); n->ival = a; n->denom = b; drop_children(n); }

static const char *builtin_name(ASTNodeType_t t);
const char *ASTNode_getName(const ASTNode_t *n)
{
  if (!n) return NULL;
  if (n->name) return n->name;
  return builtin_name(n->type);
}
long ASTNode_getInteger(const ASTNode_t *n) { return n->ival; }
double ASTNode_getReal(const ASTNode_t *n)
{
  if (n->type == AST_RATIONAL) return (double)n->ival / (double)n->denom;
  if (n->type == AST_INTEGER) return (double)n->ival;
  return n->rval;
}

int ASTNode_isInteger(const ASTNode_t *n) { return n->type == AST_INTEGER; }
int ASTNode_isReal(const ASTNode_t *n) { return n->type == AST_REAL || n->type == AST_REAL_E || n->type == AST_RATIONAL; }
int ASTNode_isNumber(const ASTNode_t *n) { return ASTNode_isInteger(n) || ASTNode_isReal(n); }
int ASTNode_isName(const ASTNode_t *n) { return n->type == AST_NAME || n->type == AST_NAME_TIME; }
int ASTNode_isOperator(const ASTNode_t *n)
{ return n->type == AST_PLUS || n->type == AST_MINUS || n->type == AST_TIMES || n->type == AST_DIVIDE || n->type == AST_POWER; }
int ASTNode_isFunction(const ASTNode_t *n) { return n->type >= AST_FUNCTION && n->type <= AST_FUNCTION_TANH; }
int ASTNode_isLogical(const ASTNode_t *n) { return n->type >= AST_LOGICAL_AND && n->type <= AST_LOGICAL_XOR; }
int ASTNode_isRelational(const ASTNode_t *n) { return n->type >= AST_RELATIONAL_EQ && n->type <= AST_RELATIONAL_NEQ; }
int ASTNode_isConstant(const ASTNode_t *n) { return n->type >= AST_CONSTANT_E && n->type <= AST_CONSTANT_TRUE; }
int ASTNode_isUMinus(const ASTNode_t *n) { return n->type == AST_MINUS && n->nchildren == 1; }
int ASTNode_isUnknown(const ASTNode_t *n) { return n->type == AST_UNKNOWN; }
int ASTNode_isLambda(const ASTNode_t *n) { return n->type == AST_LAMBDA; }
int ASTNode_isLog10(const ASTNode_t *n)
{ return n->type == AST_FUNCTION_LOG && n->nchildren == 2 && n->children[0]->type == AST_INTEGER && n->children[0]->ival == 10; }
int ASTNode_isSqrt(const ASTNode_t *n)
{ return n->type == AST_FUNCTION_ROOT && n->nchildren == 2 && n->children[0]->type == AST_INTEGER && n->children[0]->ival == 2; }

int ASTNode_isIndexName(const ASTNode_t *n) { return n->index >= 0; }
int ASTNode_isSetIndex(const ASTNode_t *n) { return n->index >= 0; }
unsigned int ASTNode_getIndex(const ASTNode_t *n) { return (unsigned int)n->index; }
void ASTNode_setIndex(ASTNode_t *n, unsigned int i) { n->index = (int)i; }
int ASTNode_isSetData(const ASTNode_t *n) { return n->index >= 0 && n->isData; }
void ASTNode_setData(ASTNode_t *n) { n->isData = 1; }

int ASTNode_getPrecedence(const ASTNode_t *n)
{
  if (ASTNode_isUMinus(n)) return 5;
  switch (n->type) {
  case AST_PLUS: case AST_MINUS: return 2;
  case AST_TIMES: case AST_DIVIDE: return 3;
  case AST_POWER: return 4;
  default: return 6;
  }
}

static void fill_list(const ASTNode_t *n, ASTNodePredicate pred, List_t *l)
{
  unsigned int i;
  if (pred(n)) List_add(l, (void *)n);
  for (i = 0; i < n->nchildren; i++) fill_list(n->children[i], pred, l);
}
List_t *ASTNode_getListOfNodes(const ASTNode_t *n, ASTNodePredicate pred)
{
  List_t *l = List_create();
  if (n) fill_list(n, pred, l);
  return l;
}

/* --------------------------------------------------- names of node kinds */
typedef struct { const char *name; ASTNodeType_t type; } name_type_t;

/* canonical (MathML-ish) names, returned by ASTNode_getName for built-ins */
static const name_type_t CANON[] = {
  {"abs", AST_FUNCTION_ABS}, {"arccos", AST_FUNCTION_ARCCOS}, {"arccosh", AST_FUNCTION_ARCCOSH},
  {"arccot", AST_FUNCTION_ARCCOT}, {"arccoth", AST_FUNCTION_ARCCOTH}, {"arccsc", AST_FUNCTION_ARCCSC},
  {"arccsch", AST_FUNCTION_ARCCSCH}, {"arcsec", AST_FUNCTION_ARCSEC}, {"arcsech", AST_FUNCTION_ARCSECH},
  {"arcsin", AST_FUNCTION_ARCSIN}, {"arcsinh", AST_FUNCTION_ARCSINH}, {"arctan", AST_FUNCTION_ARCTAN},
  {"arctanh", AST_FUNCTION_ARCTANH}, {"ceiling", AST_FUNCTION_CEILING}, {"cos", AST_FUNCTION_COS},
  {"cosh", AST_FUNCTION_COSH}, {"cot", AST_FUNCTION_COT}, {"coth", AST_FUNCTION_COTH},
  {"csc", AST_FUNCTION_CSC}, {"csch", AST_FUNCTION_CSCH}, {"delay", AST_FUNCTION_DELAY},
  {"exp", AST_FUNCTION_EXP}, {"factorial", AST_FUNCTION_FACTORIAL}, {"floor", AST_FUNCTION_FLOOR},
  {"ln", AST_FUNCTION_LN}, {"log", AST_FUNCTION_LOG}, {"piecewise", AST_FUNCTION_PIECEWISE},
  {"power", AST_FUNCTION_POWER}, {"root", AST_FUNCTION_ROOT}, {"sec", AST_FUNCTION_SEC},
  {"sech", AST_FUNCTION_SECH}, {"sin", AST_FUNCTION_SIN}, {"sinh", AST_FUNCTION_SINH},
  {"tan", AST_FUNCTION_TAN}, {"tanh", AST_FUNCTION_TANH},
  {"and", AST_LOGICAL_AND}, {"not", AST_LOGICAL_NOT}, {"or", AST_LOGICAL_OR}, {"xor", AST_LOGICAL_XOR},
  {"eq", AST_RELATIONAL_EQ}, {"geq", AST_RELATIONAL_GEQ}, {"gt", AST_RELATIONAL_GT},
  {"leq", AST_RELATIONAL_LEQ}, {"lt", AST_RELATIONAL_LT}, {"neq", AST_RELATIONAL_NEQ},
  {"exponentiale", AST_CONSTANT_E}, {"false", AST_CONSTANT_FALSE}, {"pi", AST_CONSTANT_PI},
  {"true", AST_CONSTANT_TRUE}, {"lambda", AST_LAMBDA},
  {NULL, AST_UNKNOWN}
};
/* extra spellings accepted by the Level-1 infix parser */
static const name_type_t L1_ALIASES[] = {
  {"acos", AST_FUNCTION_ARCCOS}, {"asin", AST_FUNCTION_ARCSIN}, {"atan", AST_FUNCTION_ARCTAN},
  {"ceil", AST_FUNCTION_CEILING}, {"pow", AST_FUNCTION_POWER},
  {"acosh", AST_FUNCTION_ARCCOSH}, {"asinh", AST_FUNCTION_ARCSINH}, {"atanh", AST_FUNCTION_ARCTANH},
  {NULL, AST_UNKNOWN}
};

static const char *builtin_name(ASTNodeType_t t)
{
  const name_type_t *p;
  for (p = CANON; p->name; p++) if (p->type == t) return p->name;
  return NULL;
}
static ASTNodeType_t lookup_type(const char *name)
{
  const name_type_t *p;
  for (p = CANON; p->name; p++) if (strcmp(p->name, name) == 0) return p->type;
  for (p = L1_ALIASES; p->name; p++) if (strcmp(p->name, name) == 0) return p->type;
  return AST_UNKNOWN;
}

/* ------------------------------------------------------- formula parser */
/* Precedence-climbing parser for the SBML Level 1 formula grammar:
 *   level 2: binary + -     (left)
 *   level 3: binary * /     (left)
 *   level 4: binary ^       (left)
 *   level 5: unary  -       (right)
 *   level 6: operands, function calls, ( )
 */
typedef struct { const char *s; int err; } parser_t;

static void skipws(parser_t *p) { while (*p->s && isspace((unsigned char)*p->s)) p->s++; }
static ASTNode_t *parse_expr(parser_t *p, int minprec);

static void canonicalize_function(ASTNode_t *n)
{
  /* libSBML ASTNode::canonicalizeFunctionL1 / canonicalizeFunction */
  const char *nm = n->name;
  ASTNodeType_t t;
  if (!nm) return;
  if (strcmp(nm, "log") == 0 && n->nchildren == 1) { ASTNode_setType(n, AST_FUNCTION_LN); return; }
  if (strcmp(nm, "log10") == 0 && n->nchildren == 1) {
    ASTNode_t *ten = ASTNode_create(); ASTNode_setInteger(ten, 10);
    ASTNode_setType(n, AST_FUNCTION_LOG); ASTNode_prependChild(n, ten); return;
  }
  if (strcmp(nm, "sqr") == 0 && n->nchildren == 1) {
    ASTNode_t *two = ASTNode_create(); ASTNode_setInteger(two, 2);
    ASTNode_setType(n, AST_FUNCTION_POWER); ASTNode_addChild(n, two); return;
  }
  if (strcmp(nm, "sqrt") == 0 && n->nchildren == 1) {
    ASTNode_t *two = ASTNode_create(); ASTNode_setInteger(two, 2);
    ASTNode_setType(n, AST_FUNCTION_ROOT); ASTNode_prependChild(n, two); return;
  }
  t = lookup_type(nm);
  if (t != AST_UNKNOWN && !(t >= AST_CONSTANT_E && t <= AST_CONSTANT_TRUE)) {
    if (t == AST_FUNCTION_LOG && n->nchildren == 1) {
      ASTNode_t *ten = ASTNode_create(); ASTNode_setInteger(ten, 10);
      ASTNode_prependChild(n, ten);
    }
    if (t == AST_FUNCTION_ROOT && n->nchildren == 1) {
      /* root(x): the degree 2 is made explicit for evaluation and differentiation; libSBML keeps the single argument, and
         the text generator of the reference prints it that way -- the unused rational denominator of the node remembers it */
      ASTNode_t *two = ASTNode_create(); ASTNode_setInteger(two, 2);
      ASTNode_prependChild(n, two);
      n->denom = -2;
    }
    ASTNode_setType(n, t);
  }
}

static ASTNode_t *parse_number(parser_t *p)
{
  const char *b = p->s;
  char *end;
  int isreal = 0, isexp = 0;
  ASTNode_t *n = ASTNode_create();
  while (isdigit((unsigned char)*p->s)) p->s++;
  if (*p->s == '.') { isreal = 1; p->s++; while (isdigit((unsigned char)*p->s)) p->s++; }
  if (*p->s == 'e' || *p->s == 'E') {
    const char *q = p->s + 1;
    if (*q == '+' || *q == '-') q++;
    if (isdigit((unsigned char)*q)) { isexp = 1; p->s = q; while (isdigit((unsigned char)*p->s)) p->s++; }
  }
  if (isexp) {
    double v = strtod(b, &end);
    /* keep value exactly as written; exponent kept only for the printer */
    const char *e = b; long ex;
    while (*e != 'e' && *e != 'E') e++;
    ex = strtol(e + 1, NULL, 10);
    ASTNode_setType(n, AST_REAL_E); n->rval = v; n->ival = ex;
  } else if (isreal) {
    ASTNode_setReal(n, strtod(b, &end));
  } else {
    ASTNode_setInteger(n, strtol(b, &end, 10));
  }
  return n;
}

static ASTNode_t *parse_primary(parser_t *p)
{
  ASTNode_t *n;
  skipws(p);
  if (*p->s == '(') {
    p->s++;
    n = parse_expr(p, 2);
    skipws(p);
    if (*p->s == ')') p->s++; else p->err = 1;
    return n;
  }
  if (isdigit((unsigned char)*p->s) || (*p->s == '.' && isdigit((unsigned char)p->s[1])))
    return parse_number(p);
  if (isalpha((unsigned char)*p->s) || *p->s == '_') {
    const char *b = p->s;
    char *name;
    size_t len;
    while (isalnum((unsigned char)*p->s) || *p->s == '_') p->s++;
    len = (size_t)(p->s - b);
    name = (char *)malloc(len + 1); memcpy(name, b, len); name[len] = 0;
    skipws(p);
    n = ASTNode_create();
    if (*p->s == '(') {                      /* function call */
      p->s++;
      n->type = AST_FUNCTION; n->name = name;
      skipws(p);
      if (*p->s == ')') p->s++;
      else for (;;) {
        ASTNode_addChild(n, parse_expr(p, 2));
        skipws(p);
        if (*p->s == ',') { p->s++; continue; }
        if (*p->s == ')') { p->s++; break; }
        p->err = 1; break;
      }
      canonicalize_function(n);
    } else {                                 /* symbol */
      ASTNodeType_t t = lookup_type(name);
      if (t >= AST_CONSTANT_E && t <= AST_CONSTANT_TRUE) { n->type = t; free(name); }
      else { n->type = AST_NAME; n->name = name; }
    }
    return n;
  }
  p->err = 1;
  return ASTNode_create();
}

static ASTNode_t *parse_unary(parser_t *p)
{
  skipws(p);
  if (*p->s == '-') {
    ASTNode_t *c, *n;
    p->s++;
    c = parse_unary(p);
    /* libSBML FormulaParser folds "- <number>" into a negative number */
    if (c->type == AST_INTEGER) { c->ival = -c->ival; return c; }
    if (c->type == AST_REAL || c->type == AST_REAL_E) { c->rval = -c->rval; return c; }
    n = ASTNode_createWithType(AST_MINUS);
    ASTNode_addChild(n, c);
    return n;
  }
  return parse_primary(p);
}

static int binprec(char c) { return (c == '+' || c == '-') ? 2 : (c == '*' || c == '/') ? 3 : (c == '^') ? 4 : 0; }

static ASTNode_t *parse_expr(parser_t *p, int minprec)
{
  ASTNode_t *lhs = parse_unary(p);
  for (;;) {
    char op; int prec; ASTNode_t *rhs, *n;
    skipws(p);
    op = *p->s; prec = binprec(op);
    if (!prec || prec < minprec) break;
    p->s++;
    rhs = parse_expr(p, prec + 1);            /* all binary operators associate left */
    n = ASTNode_createWithType((ASTNodeType_t)op);
    ASTNode_addChild(n, lhs); ASTNode_addChild(n, rhs);
    lhs = n;
  }
  return lhs;
}

ASTNode_t *SBML_parseFormula(const char *formula)
{
  parser_t p; ASTNode_t *n;
  if (!formula) return NULL;
  p.s = formula; p.err = 0;
  n = parse_expr(&p, 2);
  skipws(&p);
  if (p.err || *p.s) { ASTNode_free(n); return NULL; }
  return n;
}

/* ------------------------------------------------------ formula printer */
typedef struct { char *b; size_t len, cap; } sbuf_t;
static void sb_put(sbuf_t *s, const char *t)
{
  size_t n = strlen(t);
  if (s->len + n + 1 > s->cap) { s->cap = 2 * (s->len + n + 1); s->b = (char *)realloc(s->b, s->cap); }
  memcpy(s->b + s->len, t, n + 1); s->len += n;
}

static int fmt_is_function(const ASTNode_t *n)
{ return ASTNode_isFunction(n) || ASTNode_isLambda(n) || ASTNode_isLogical(n) || ASTNode_isRelational(n); }

static int fmt_is_grouped(const ASTNode_t *parent, const ASTNode_t *child)
{
  int pp, cp, group = 0;
  if (parent && !fmt_is_function(parent)) {
    group = 1;
    pp = ASTNode_getPrecedence(parent); cp = ASTNode_getPrecedence(child);
    if (pp < cp) group = 0;
    else if (pp == cp) {
      if (ASTNode_getRightChild(parent) == child) {
        ASTNodeType_t pt = parent->type, ct = child->type;
        group = ((pt == AST_MINUS || pt == AST_DIVIDE) || (pt != ct));
      } else group = 0;
    }
  }
  return group;
}

static void fmt_visit(const ASTNode_t *parent, const ASTNode_t *n, sbuf_t *s);

static void fmt_real(const ASTNode_t *n, sbuf_t *s)
{
  char buf[64];
  double v = ASTNode_getReal(n);
  if (isnan(v)) sb_put(s, "NaN");
  else if (isinf(v)) sb_put(s, v < 0 ? "-INF" : "INF");
  else if (v == 0.0 && signbit(v)) sb_put(s, "-0");
  else {
    if (n->type == AST_REAL_E) snprintf(buf, sizeof buf, "%e", v);
    else snprintf(buf, sizeof buf, "%.15g", v);
    sb_put(s, buf);
  }
}

static const char *fmt_function_name(const ASTNode_t *n)
{
  switch (n->type) {
  case AST_FUNCTION_ARCCOS: return "acos";
  case AST_FUNCTION_ARCSIN: return "asin";
  case AST_FUNCTION_ARCTAN: return "atan";
  case AST_FUNCTION_CEILING: return "ceil";
  case AST_FUNCTION_LN: return "log";
  case AST_FUNCTION_POWER: return "pow";
  default: return ASTNode_getName(n);
  }
}

static void fmt_args(const ASTNode_t *n, unsigned int from, sbuf_t *s)
{
  unsigned int i;
  sb_put(s, "(");
  for (i = from; i < n->nchildren; i++) {
    if (i > from) sb_put(s, ", ");
    fmt_visit(n, n->children[i], s);
  }
  sb_put(s, ")");
}

static void fmt_visit(const ASTNode_t *parent, const ASTNode_t *n, sbuf_t *s)
{
  char buf[64];
  int grouped;
  if (ASTNode_isLog10(n)) { sb_put(s, "log10"); fmt_args(n, 1, s); return; }
  if (ASTNode_isSqrt(n)) { sb_put(s, "sqrt"); fmt_args(n, 1, s); return; }
  if (fmt_is_function(n)) { const char *nm = fmt_function_name(n); sb_put(s, nm ? nm : "unknown"); fmt_args(n, 0, s); return; }
  if (ASTNode_isUMinus(n)) {
    /* libSBML FormulaFormatter_visitUMinus: prints "-" then the operand */
    sb_put(s, "-");
    fmt_visit(n, n->children[0], s);
    return;
  }
  grouped = fmt_is_grouped(parent, n);
  if (grouped) sb_put(s, "(");
  switch (n->type) {
  case AST_PLUS: case AST_MINUS: case AST_TIMES: case AST_DIVIDE: case AST_POWER:
    if (n->nchildren > 0) fmt_visit(n, n->children[0], s);
    if (n->type == AST_POWER) { buf[0] = '^'; buf[1] = 0; }
    else { buf[0] = ' '; buf[1] = (char)n->type; buf[2] = ' '; buf[3] = 0; }
    sb_put(s, buf);
    if (n->nchildren > 1) fmt_visit(n, n->children[n->nchildren - 1], s);
    break;
  case AST_INTEGER: snprintf(buf, sizeof buf, "%ld", n->ival); sb_put(s, buf); break;
  case AST_RATIONAL: snprintf(buf, sizeof buf, "(%ld/%ld)", n->ival, n->denom); sb_put(s, buf); break;
  case AST_REAL: case AST_REAL_E: fmt_real(n, s); break;
  default: {
      const char *nm = ASTNode_getName(n);
      sb_put(s, nm ? nm : "unknown");
    }
  }
  if (grouped) sb_put(s, ")");
}

char *SBML_formulaToString(const ASTNode_t *n)
{
  sbuf_t s; s.b = NULL; s.len = 0; s.cap = 0;
  sb_put(&s, "");
  if (n) fmt_visit(NULL, n, &s);
  return s.b;
}
