/*
 * sbml_io.c -- SBML (L1/L2) + MathML reader and writer on expat, and the
 * object-model functions declared in sbml_model.h.
 *
 * Role in the reference: src/sbml.c:68-118 (parseModel: read, L1->L2 convert)
 * delegates all of this to libSBML.  Here an element tree is built with expat
 * and walked.  Level 1 documents are converted while reading exactly like the
 * reference's forced L1->L2 conversion: `name` becomes the id, `volume` the
 * size, formula strings are parsed with the Level-1 infix grammar, L1 rule
 * kinds map to assignment/rate rules.
 */
#include <expat.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sbmlsolver/sbmlModel.h"

/* ------------------------------------------------------------ tiny DOM */
typedef struct xnode {
  char *name;
  char **attr;              /* name,value,... NULL */
  char *text;
  struct xnode **kids; unsigned int nkids, cap;
  struct xnode *parent;
} xnode_t;

static char *xstrdup(const char *s) { char *r; if (!s) return NULL; r = (char *)malloc(strlen(s) + 1); strcpy(r, s); return r; }

static const char *localname(const char *qname)
{
  const char *c = strrchr(qname, ':');
  return c ? c + 1 : qname;
}
static void xfree(xnode_t *n)
{
  unsigned int i;
  if (!n) return;
  for (i = 0; i < n->nkids; i++) xfree(n->kids[i]);
  free(n->kids); free(n->name); free(n->text);
  if (n->attr) { for (i = 0; n->attr[i]; i++) free(n->attr[i]); free(n->attr); }
  free(n);
}
typedef struct { xnode_t *root, *cur; } xbuild_t;

static void XMLCALL x_start(void *ud, const XML_Char *name, const XML_Char **atts)
{
  xbuild_t *b = (xbuild_t *)ud;
  xnode_t *n = (xnode_t *)calloc(1, sizeof(xnode_t));
  unsigned int na = 0, i;
  n->name = xstrdup(localname(name));
  while (atts[na]) na++;
  n->attr = (char **)calloc(na + 1, sizeof(char *));
  for (i = 0; i < na; i++) n->attr[i] = xstrdup((i % 2 == 0) ? localname(atts[i]) : atts[i]);
  n->parent = b->cur;
  if (b->cur) {
    xnode_t *p = b->cur;
    if (p->nkids == p->cap) { p->cap = p->cap ? 2 * p->cap : 4; p->kids = (xnode_t **)realloc(p->kids, p->cap * sizeof(xnode_t *)); }
    p->kids[p->nkids++] = n;
  } else b->root = n;
  b->cur = n;
}
static void XMLCALL x_end(void *ud, const XML_Char *name) { xbuild_t *b = (xbuild_t *)ud; (void)name; b->cur = b->cur->parent; }
static void XMLCALL x_text(void *ud, const XML_Char *s, int len)
{
  xbuild_t *b = (xbuild_t *)ud; xnode_t *n = b->cur; size_t old;
  if (!n) return;
  old = n->text ? strlen(n->text) : 0;
  n->text = (char *)realloc(n->text, old + (size_t)len + 1);
  memcpy(n->text + old, s, (size_t)len); n->text[old + (size_t)len] = 0;
}
static xnode_t *xparse(const char *xml)
{
  xbuild_t b; XML_Parser p = XML_ParserCreate(NULL);
  b.root = b.cur = NULL;
  XML_SetUserData(p, &b);
  XML_SetElementHandler(p, x_start, x_end);
  XML_SetCharacterDataHandler(p, x_text);
  if (XML_Parse(p, xml, (int)strlen(xml), 1) == XML_STATUS_ERROR) { XML_ParserFree(p); xfree(b.root); return NULL; }
  XML_ParserFree(p);
  return b.root;
}
static const char *xattr(const xnode_t *n, const char *name)
{
  unsigned int i;
  for (i = 0; n->attr && n->attr[i]; i += 2) if (strcmp(n->attr[i], name) == 0) return n->attr[i + 1];
  return NULL;
}
static xnode_t *xchild(const xnode_t *n, const char *name)
{
  unsigned int i;
  for (i = 0; i < n->nkids; i++) if (strcmp(n->kids[i]->name, name) == 0) return n->kids[i];
  return NULL;
}
static char *trimmed(const char *s)
{
  const char *b, *e; char *r;
  if (!s) return xstrdup("");
  b = s; while (*b == ' ' || *b == '\n' || *b == '\t' || *b == '\r') b++;
  e = b + strlen(b); while (e > b && (e[-1] == ' ' || e[-1] == '\n' || e[-1] == '\t' || e[-1] == '\r')) e--;
  r = (char *)malloc((size_t)(e - b) + 1); memcpy(r, b, (size_t)(e - b)); r[e - b] = 0;
  return r;
}
static int xbool(const char *v, int dflt) { if (!v) return dflt; return (strcmp(v, "true") == 0 || strcmp(v, "1") == 0); }
static double xdouble(const char *v)
{
  if (!v) return 0.0;
  if (strcmp(v, "INF") == 0) return INFINITY;
  if (strcmp(v, "-INF") == 0) return -INFINITY;
  if (strcmp(v, "NaN") == 0) return NAN;
  return strtod(v, NULL);
}

/* ------------------------------------------------------ MathML -> AST */
typedef struct { const char *tag; ASTNodeType_t type; } tag_type_t;
static const tag_type_t MATHML_OPS[] = {
  {"plus", AST_PLUS}, {"minus", AST_MINUS}, {"times", AST_TIMES}, {"divide", AST_DIVIDE},
  {"power", AST_POWER},
  {"abs", AST_FUNCTION_ABS}, {"arccos", AST_FUNCTION_ARCCOS}, {"arccosh", AST_FUNCTION_ARCCOSH},
  {"arccot", AST_FUNCTION_ARCCOT}, {"arccoth", AST_FUNCTION_ARCCOTH}, {"arccsc", AST_FUNCTION_ARCCSC},
  {"arccsch", AST_FUNCTION_ARCCSCH}, {"arcsec", AST_FUNCTION_ARCSEC}, {"arcsech", AST_FUNCTION_ARCSECH},
  {"arcsin", AST_FUNCTION_ARCSIN}, {"arcsinh", AST_FUNCTION_ARCSINH}, {"arctan", AST_FUNCTION_ARCTAN},
  {"arctanh", AST_FUNCTION_ARCTANH}, {"ceiling", AST_FUNCTION_CEILING}, {"cos", AST_FUNCTION_COS},
  {"cosh", AST_FUNCTION_COSH}, {"cot", AST_FUNCTION_COT}, {"coth", AST_FUNCTION_COTH},
  {"csc", AST_FUNCTION_CSC}, {"csch", AST_FUNCTION_CSCH}, {"exp", AST_FUNCTION_EXP},
  {"factorial", AST_FUNCTION_FACTORIAL}, {"floor", AST_FUNCTION_FLOOR}, {"ln", AST_FUNCTION_LN},
  {"log", AST_FUNCTION_LOG}, {"root", AST_FUNCTION_ROOT}, {"sec", AST_FUNCTION_SEC},
  {"sech", AST_FUNCTION_SECH}, {"sin", AST_FUNCTION_SIN}, {"sinh", AST_FUNCTION_SINH},
  {"tan", AST_FUNCTION_TAN}, {"tanh", AST_FUNCTION_TANH},
  {"and", AST_LOGICAL_AND}, {"not", AST_LOGICAL_NOT}, {"or", AST_LOGICAL_OR}, {"xor", AST_LOGICAL_XOR},
  {"eq", AST_RELATIONAL_EQ}, {"geq", AST_RELATIONAL_GEQ}, {"gt", AST_RELATIONAL_GT},
  {"leq", AST_RELATIONAL_LEQ}, {"lt", AST_RELATIONAL_LT}, {"neq", AST_RELATIONAL_NEQ},
  {NULL, AST_UNKNOWN}
};

static ASTNode_t *mathml_to_ast(const xnode_t *x);

static ASTNode_t *mathml_cn(const xnode_t *x)
{
  ASTNode_t *n = ASTNode_create();
  const char *type = xattr(x, "type");
  char *t = trimmed(x->text);
  xnode_t *sep = xchild(x, "sep");
  if (type && strcmp(type, "integer") == 0) ASTNode_setInteger(n, strtol(t, NULL, 10));
  else if (type && (strcmp(type, "e-notation") == 0 || strcmp(type, "rational") == 0) && sep) {
    /* text of <cn> a <sep/> b </cn>: expat delivers "a" and "b" concatenated around the
       empty element; recover the two halves from the raw text pieces */
    char *a = trimmed(x->text), *b = trimmed(sep->text ? sep->text : "");
    /* text after <sep/> lands in x->text as well; split on whitespace */
    char *sp = a; double m; long e;
    while (*sp && *sp != ' ' && *sp != '\n' && *sp != '\t') sp++;
    m = strtod(a, NULL);
    e = strtol(*sp ? sp : b, NULL, 10);
    if (strcmp(type, "rational") == 0) ASTNode_setRational(n, (long)m, e);
    else ASTNode_setRealWithExponent(n, m, e);
    free(a); free(b);
  } else {
    /* libSBML: a <cn> without type holding an integer literal stays a real */
    ASTNode_setReal(n, xdouble(t));
    if (!type || strcmp(type, "real") != 0) {
      /* untyped "2" is read by libSBML as AST_REAL 2 as well; keep real */
    }
  }
  free(t);
  return n;
}

static ASTNode_t *mathml_apply(const xnode_t *x)
{
  const xnode_t *op;
  ASTNode_t *n;
  const tag_type_t *p;
  unsigned int i;
  if (x->nkids == 0) return NULL;
  op = x->kids[0];
  n = ASTNode_create();
  if (strcmp(op->name, "ci") == 0) {                  /* user function call */
    char *nm = trimmed(op->text);
    n->type = AST_FUNCTION; n->name = nm;
  } else if (strcmp(op->name, "csymbol") == 0) {     /* delay */
    char *nm = trimmed(op->text);
    n->type = AST_FUNCTION_DELAY; n->name = nm;
  } else {
    for (p = MATHML_OPS; p->tag; p++) if (strcmp(p->tag, op->name) == 0) break;
    if (!p->tag) { ASTNode_free(n); return NULL; }
    n->type = p->type;
  }
  for (i = 1; i < x->nkids; i++) {
    const xnode_t *k = x->kids[i];
    ASTNode_t *c;
    if (strcmp(k->name, "degree") == 0 || strcmp(k->name, "logbase") == 0) {
      c = k->nkids ? mathml_to_ast(k->kids[0]) : NULL;
      if (!c) { ASTNode_free(n); return NULL; }
      ASTNode_prependChild(n, c);
      continue;
    }
    c = mathml_to_ast(k);
    if (!c) { ASTNode_free(n); return NULL; }
    ASTNode_addChild(n, c);
  }
  /* defaults: sqrt and log10 */
  if (n->type == AST_FUNCTION_ROOT && n->nchildren == 1) { ASTNode_t *two = ASTNode_create(); ASTNode_setInteger(two, 2); ASTNode_prependChild(n, two); }
  if (n->type == AST_FUNCTION_LOG && n->nchildren == 1) { ASTNode_t *ten = ASTNode_create(); ASTNode_setInteger(ten, 10); ASTNode_prependChild(n, ten); }
  return n;
}

static ASTNode_t *mathml_to_ast(const xnode_t *x)
{
  ASTNode_t *n;
  unsigned int i;
  if (strcmp(x->name, "math") == 0 || strcmp(x->name, "semantics") == 0)
    return x->nkids ? mathml_to_ast(x->kids[0]) : NULL;
  if (strcmp(x->name, "apply") == 0) return mathml_apply(x);
  if (strcmp(x->name, "cn") == 0) return mathml_cn(x);
  n = ASTNode_create();
  if (strcmp(x->name, "ci") == 0) { n->type = AST_NAME; n->name = trimmed(x->text); return n; }
  if (strcmp(x->name, "csymbol") == 0) {
    const char *url = xattr(x, "definitionURL");
    n->name = trimmed(x->text);
    n->type = (url && strstr(url, "delay")) ? AST_FUNCTION_DELAY : AST_NAME_TIME;
    return n;
  }
  if (strcmp(x->name, "pi") == 0) { n->type = AST_CONSTANT_PI; return n; }
  if (strcmp(x->name, "exponentiale") == 0) { n->type = AST_CONSTANT_E; return n; }
  if (strcmp(x->name, "true") == 0) { n->type = AST_CONSTANT_TRUE; return n; }
  if (strcmp(x->name, "false") == 0) { n->type = AST_CONSTANT_FALSE; return n; }
  if (strcmp(x->name, "infinity") == 0) { ASTNode_setReal(n, INFINITY); return n; }
  if (strcmp(x->name, "notanumber") == 0) { ASTNode_setReal(n, NAN); return n; }
  if (strcmp(x->name, "piecewise") == 0) {
    n->type = AST_FUNCTION_PIECEWISE;
    for (i = 0; i < x->nkids; i++) {
      const xnode_t *k = x->kids[i]; unsigned int j;
      for (j = 0; j < k->nkids; j++) {
        ASTNode_t *c = mathml_to_ast(k->kids[j]);
        if (!c) { ASTNode_free(n); return NULL; }
        ASTNode_addChild(n, c);
      }
    }
    return n;
  }
  if (strcmp(x->name, "lambda") == 0) {
    n->type = AST_LAMBDA;
    for (i = 0; i < x->nkids; i++) {
      const xnode_t *k = x->kids[i];
      ASTNode_t *c = (strcmp(k->name, "bvar") == 0) ? (k->nkids ? mathml_to_ast(k->kids[0]) : NULL) : mathml_to_ast(k);
      if (!c) { ASTNode_free(n); return NULL; }
      ASTNode_addChild(n, c);
    }
    return n;
  }
  ASTNode_free(n);
  return NULL;
}

static ASTNode_t *math_from_xnode(const xnode_t *math)
{
  ASTNode_t *n = math ? mathml_to_ast(math) : NULL;
  /* libSBML hands n-ary plus/times to the reference as left-nested binary trees */
  if (n) ASTNode_reduceToBinary(n);
  return n;
}

ASTNode_t *readMathMLFromString(const char *xml)
{
  xnode_t *root = xparse(xml);
  ASTNode_t *n;
  if (!root) return NULL;
  n = math_from_xnode(root);
  xfree(root);
  return n;
}

/* ------------------------------------------------------ AST -> MathML */
typedef struct { char *b; size_t len, cap; } wbuf_t;
static void wput(wbuf_t *w, const char *t)
{
  size_t n = strlen(t);
  if (w->len + n + 1 > w->cap) { w->cap = 2 * (w->len + n + 1); w->b = (char *)realloc(w->b, w->cap); }
  memcpy(w->b + w->len, t, n + 1); w->len += n;
}
static void wprintf_(wbuf_t *w, const char *fmt, ...);
#include <stdarg.h>
static void wprintf_(wbuf_t *w, const char *fmt, ...)
{
  char buf[1024]; va_list ap;
  va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  wput(w, buf);
}
static const char *mathml_tag(ASTNodeType_t t)
{
  const tag_type_t *p;
  if (t == AST_FUNCTION_POWER) return "power";
  for (p = MATHML_OPS; p->tag; p++) if (p->type == t) return p->tag;
  return NULL;
}
static void ast_to_mathml(const ASTNode_t *n, wbuf_t *w)
{
  unsigned int i;
  switch (n->type) {
  case AST_INTEGER: wprintf_(w, "<cn type=\"integer\">%ld</cn>", n->ival); return;
  case AST_REAL: case AST_REAL_E: wprintf_(w, "<cn>%.17g</cn>", n->rval); return;
  case AST_RATIONAL: wprintf_(w, "<cn type=\"rational\">%ld<sep/>%ld</cn>", n->ival, n->denom); return;
  case AST_NAME: wprintf_(w, "<ci>%s</ci>", n->name ? n->name : ""); return;
  case AST_NAME_TIME:
    wprintf_(w, "<csymbol encoding=\"text\" definitionURL=\"http://www.sbml.org/sbml/symbols/time\">%s</csymbol>", n->name ? n->name : "time"); return;
  case AST_CONSTANT_E: wput(w, "<exponentiale/>"); return;
  case AST_CONSTANT_PI: wput(w, "<pi/>"); return;
  case AST_CONSTANT_TRUE: wput(w, "<true/>"); return;
  case AST_CONSTANT_FALSE: wput(w, "<false/>"); return;
  case AST_FUNCTION_PIECEWISE:
    wput(w, "<piecewise>");
    for (i = 0; i + 1 < n->nchildren; i += 2) {
      wput(w, "<piece>"); ast_to_mathml(n->children[i], w); ast_to_mathml(n->children[i + 1], w); wput(w, "</piece>");
    }
    if (i < n->nchildren) { wput(w, "<otherwise>"); ast_to_mathml(n->children[i], w); wput(w, "</otherwise>"); }
    wput(w, "</piecewise>"); return;
  case AST_LAMBDA:
    wput(w, "<lambda>");
    for (i = 0; i + 1 < n->nchildren; i++) { wput(w, "<bvar>"); ast_to_mathml(n->children[i], w); wput(w, "</bvar>"); }
    if (n->nchildren) ast_to_mathml(n->children[n->nchildren - 1], w);
    wput(w, "</lambda>"); return;
  default: break;
  }
  wput(w, "<apply>");
  if (n->type == AST_FUNCTION) wprintf_(w, "<ci>%s</ci>", n->name ? n->name : "");
  else if (n->type == AST_FUNCTION_DELAY)
    wprintf_(w, "<csymbol encoding=\"text\" definitionURL=\"http://www.sbml.org/sbml/symbols/delay\">%s</csymbol>", n->name ? n->name : "delay");
  else wprintf_(w, "<%s/>", mathml_tag(n->type) ? mathml_tag(n->type) : "unknown");
  for (i = 0; i < n->nchildren; i++) {
    if (i == 0 && n->nchildren == 2 && n->type == AST_FUNCTION_ROOT) { wput(w, "<degree>"); ast_to_mathml(n->children[0], w); wput(w, "</degree>"); continue; }
    if (i == 0 && n->nchildren == 2 && n->type == AST_FUNCTION_LOG) { wput(w, "<logbase>"); ast_to_mathml(n->children[0], w); wput(w, "</logbase>"); continue; }
    ast_to_mathml(n->children[i], w);
  }
  wput(w, "</apply>");
}
char *writeMathMLToString(const ASTNode_t *n)
{
  wbuf_t w; w.b = NULL; w.len = w.cap = 0;
  wput(&w, "<math xmlns=\"http://www.w3.org/1998/Math/MathML\">");
  if (n) ast_to_mathml(n, &w);
  wput(&w, "</math>");
  return w.b;
}

/* ------------------------------------------------------- object model */
void ListOf_append(ListOf_t *l, void *p)
{
  if (l->size == l->cap) { l->cap = l->cap ? 2 * l->cap : 8; l->items = (void **)realloc(l->items, l->cap * sizeof(void *)); }
  l->items[l->size++] = p;
}
unsigned int ListOf_size(const ListOf_t *l) { return l ? l->size : 0; }
void *ListOf_get(const ListOf_t *l, unsigned int i) { return (l && i < l->size) ? l->items[i] : NULL; }

Model_t *Model_create(unsigned int level, unsigned int version)
{
  Model_t *m = (Model_t *)calloc(1, sizeof(Model_t));
  m->level = level; m->version = version;
  return m;
}
void Model_setId(Model_t *m, const char *id) { free(m->id); m->id = xstrdup(id); }
const char *Model_getId(const Model_t *m) { return m->id; }

Compartment_t *Model_createCompartment(Model_t *m, const char *id, double size)
{
  Compartment_t *c = (Compartment_t *)calloc(1, sizeof(Compartment_t));
  c->id = xstrdup(id); c->size = size; c->isSetSize = 1; c->spatialDimensions = 3; c->constant = 1;
  ListOf_append(&m->compartments, c);
  return c;
}
Species_t *Model_createSpecies(Model_t *m, const char *id, const char *comp, double conc)
{
  Species_t *s = (Species_t *)calloc(1, sizeof(Species_t));
  s->id = xstrdup(id); s->compartment = xstrdup(comp);
  s->initialConcentration = conc; s->isSetInitialConcentration = 1;
  ListOf_append(&m->species, s);
  return s;
}
Parameter_t *Model_createParameter(Model_t *m, const char *id, double value)
{
  Parameter_t *p = (Parameter_t *)calloc(1, sizeof(Parameter_t));
  p->id = xstrdup(id); p->value = value; p->isSetValue = 1; p->constant = 1;
  ListOf_append(&m->parameters, p);
  return p;
}
Reaction_t *Model_createReaction(Model_t *m, const char *id)
{
  Reaction_t *r = (Reaction_t *)calloc(1, sizeof(Reaction_t));
  r->id = xstrdup(id); r->reversible = 1;
  ListOf_append(&m->reactions, r);
  return r;
}
static SpeciesReference_t *sref_new(const char *species, double st)
{
  SpeciesReference_t *s = (SpeciesReference_t *)calloc(1, sizeof(SpeciesReference_t));
  s->species = xstrdup(species); s->stoichiometry = st;
  return s;
}
SpeciesReference_t *Reaction_createReactant(Reaction_t *r, const char *sp, double st) { SpeciesReference_t *s = sref_new(sp, st); ListOf_append(&r->reactants, s); return s; }
SpeciesReference_t *Reaction_createProduct(Reaction_t *r, const char *sp, double st) { SpeciesReference_t *s = sref_new(sp, st); ListOf_append(&r->products, s); return s; }
SpeciesReference_t *Reaction_createModifier(Reaction_t *r, const char *sp) { SpeciesReference_t *s = sref_new(sp, 1.0); ListOf_append(&r->modifiers, s); return s; }
KineticLaw_t *Reaction_createKineticLaw(Reaction_t *r, const char *formula)
{
  KineticLaw_t *k = (KineticLaw_t *)calloc(1, sizeof(KineticLaw_t));
  k->math = formula ? SBML_parseFormula(formula) : NULL;
  r->kineticLaw = k;
  return k;
}
Parameter_t *KineticLaw_createParameter(KineticLaw_t *k, const char *id, double value)
{
  Parameter_t *p = (Parameter_t *)calloc(1, sizeof(Parameter_t));
  p->id = xstrdup(id); p->value = value; p->isSetValue = 1; p->constant = 1;
  ListOf_append(&k->parameters, p);
  return p;
}
Rule_t *Model_createRule(Model_t *m, SBMLTypeCode_t type, const char *variable, ASTNode_t *math)
{
  Rule_t *r = (Rule_t *)calloc(1, sizeof(Rule_t));
  r->type = type; r->variable = xstrdup(variable); r->math = math;
  ListOf_append(&m->rules, r);
  return r;
}
Rule_t *Model_createRuleFromFormula(Model_t *m, SBMLTypeCode_t type, const char *variable, const char *formula)
{ return Model_createRule(m, type, variable, SBML_parseFormula(formula)); }
Event_t *Model_createEvent(Model_t *m, const char *id, ASTNode_t *trigger)
{
  Event_t *e = (Event_t *)calloc(1, sizeof(Event_t));
  e->id = xstrdup(id); e->trigger = trigger;
  ListOf_append(&m->events, e);
  return e;
}
EventAssignment_t *Event_createEventAssignment(Event_t *e, const char *variable, ASTNode_t *math)
{
  EventAssignment_t *a = (EventAssignment_t *)calloc(1, sizeof(EventAssignment_t));
  a->variable = xstrdup(variable); a->math = math;
  ListOf_append(&e->assignments, a);
  return a;
}
InitialAssignment_t *Model_createInitialAssignment(Model_t *m, const char *symbol, ASTNode_t *math)
{
  InitialAssignment_t *a = (InitialAssignment_t *)calloc(1, sizeof(InitialAssignment_t));
  a->symbol = xstrdup(symbol); a->math = math;
  ListOf_append(&m->initialAssignments, a);
  return a;
}
FunctionDefinition_t *Model_createFunctionDefinition(Model_t *m, const char *id, ASTNode_t *lambda)
{
  FunctionDefinition_t *f = (FunctionDefinition_t *)calloc(1, sizeof(FunctionDefinition_t));
  f->id = xstrdup(id); f->math = lambda;
  ListOf_append(&m->functionDefinitions, f);
  return f;
}

unsigned int Model_getNumCompartments(const Model_t *m) { return m->compartments.size; }
unsigned int Model_getNumSpecies(const Model_t *m) { return m->species.size; }
unsigned int Model_getNumParameters(const Model_t *m) { return m->parameters.size; }
unsigned int Model_getNumReactions(const Model_t *m) { return m->reactions.size; }
unsigned int Model_getNumRules(const Model_t *m) { return m->rules.size; }
unsigned int Model_getNumEvents(const Model_t *m) { return m->events.size; }
unsigned int Model_getNumInitialAssignments(const Model_t *m) { return m->initialAssignments.size; }
unsigned int Model_getNumFunctionDefinitions(const Model_t *m) { return m->functionDefinitions.size; }
Compartment_t *Model_getCompartment(const Model_t *m, unsigned int i) { return (Compartment_t *)ListOf_get(&m->compartments, i); }
Species_t *Model_getSpecies(const Model_t *m, unsigned int i) { return (Species_t *)ListOf_get(&m->species, i); }
Parameter_t *Model_getParameter(const Model_t *m, unsigned int i) { return (Parameter_t *)ListOf_get(&m->parameters, i); }
Reaction_t *Model_getReaction(const Model_t *m, unsigned int i) { return (Reaction_t *)ListOf_get(&m->reactions, i); }
Rule_t *Model_getRule(const Model_t *m, unsigned int i) { return (Rule_t *)ListOf_get(&m->rules, i); }
Event_t *Model_getEvent(const Model_t *m, unsigned int i) { return (Event_t *)ListOf_get(&m->events, i); }
InitialAssignment_t *Model_getInitialAssignment(const Model_t *m, unsigned int i) { return (InitialAssignment_t *)ListOf_get(&m->initialAssignments, i); }
FunctionDefinition_t *Model_getFunctionDefinition(const Model_t *m, unsigned int i) { return (FunctionDefinition_t *)ListOf_get(&m->functionDefinitions, i); }

Compartment_t *Model_getCompartmentById(const Model_t *m, const char *id)
{ unsigned int i; if (!id) return NULL; for (i = 0; i < m->compartments.size; i++) { Compartment_t *c = Model_getCompartment(m, i); if (c->id && strcmp(c->id, id) == 0) return c; } return NULL; }
Species_t *Model_getSpeciesById(const Model_t *m, const char *id)
{ unsigned int i; if (!id) return NULL; for (i = 0; i < m->species.size; i++) { Species_t *s = Model_getSpecies(m, i); if (s->id && strcmp(s->id, id) == 0) return s; } return NULL; }
Parameter_t *Model_getParameterById(const Model_t *m, const char *id)
{ unsigned int i; if (!id) return NULL; for (i = 0; i < m->parameters.size; i++) { Parameter_t *p = Model_getParameter(m, i); if (p->id && strcmp(p->id, id) == 0) return p; } return NULL; }
Reaction_t *Model_getReactionById(const Model_t *m, const char *id)
{ unsigned int i; if (!id) return NULL; for (i = 0; i < m->reactions.size; i++) { Reaction_t *r = Model_getReaction(m, i); if (r->id && strcmp(r->id, id) == 0) return r; } return NULL; }

Compartment_t *Compartment_clone(const Compartment_t *c) { Compartment_t *n = (Compartment_t *)malloc(sizeof *n); *n = *c; n->id = xstrdup(c->id); return n; }
Species_t *Species_clone(const Species_t *s) { Species_t *n = (Species_t *)malloc(sizeof *n); *n = *s; n->id = xstrdup(s->id); n->compartment = xstrdup(s->compartment); return n; }
Parameter_t *Parameter_clone(const Parameter_t *p) { Parameter_t *n = (Parameter_t *)malloc(sizeof *n); *n = *p; n->id = xstrdup(p->id); return n; }
void Parameter_setId(Parameter_t *p, const char *id) { free(p->id); p->id = xstrdup(id); }
Event_t *Event_clone(const Event_t *e)
{
  unsigned int i;
  Event_t *n = (Event_t *)calloc(1, sizeof *n);
  n->id = xstrdup(e->id); n->trigger = ASTNode_deepCopy(e->trigger); n->delay = ASTNode_deepCopy(e->delay);
  for (i = 0; i < e->assignments.size; i++) {
    EventAssignment_t *a = (EventAssignment_t *)e->assignments.items[i];
    Event_createEventAssignment(n, a->variable, ASTNode_deepCopy(a->math));
  }
  return n;
}

static SpeciesReference_t *sref_clone(const SpeciesReference_t *s)
{ SpeciesReference_t *n = sref_new(s->species, s->stoichiometry); n->stoichiometryMath = ASTNode_deepCopy(s->stoichiometryMath); return n; }

Model_t *Model_clone(const Model_t *m)
{
  unsigned int i, j;
  Model_t *n = Model_create(m->level, m->version);
  n->id = xstrdup(m->id); n->name = xstrdup(m->name); n->notes = xstrdup(m->notes);
  for (i = 0; i < m->functionDefinitions.size; i++) { FunctionDefinition_t *f = Model_getFunctionDefinition(m, i); Model_createFunctionDefinition(n, f->id, ASTNode_deepCopy(f->math)); }
  for (i = 0; i < m->compartments.size; i++) ListOf_append(&n->compartments, Compartment_clone(Model_getCompartment(m, i)));
  for (i = 0; i < m->species.size; i++) ListOf_append(&n->species, Species_clone(Model_getSpecies(m, i)));
  for (i = 0; i < m->parameters.size; i++) ListOf_append(&n->parameters, Parameter_clone(Model_getParameter(m, i)));
  for (i = 0; i < m->initialAssignments.size; i++) { InitialAssignment_t *a = Model_getInitialAssignment(m, i); Model_createInitialAssignment(n, a->symbol, ASTNode_deepCopy(a->math)); }
  for (i = 0; i < m->rules.size; i++) { Rule_t *r = Model_getRule(m, i); Model_createRule(n, r->type, r->variable, ASTNode_deepCopy(r->math)); }
  for (i = 0; i < m->reactions.size; i++) {
    Reaction_t *r = Model_getReaction(m, i), *q = Model_createReaction(n, r->id);
    q->reversible = r->reversible; q->fast = r->fast;
    for (j = 0; j < r->reactants.size; j++) ListOf_append(&q->reactants, sref_clone((SpeciesReference_t *)r->reactants.items[j]));
    for (j = 0; j < r->products.size; j++) ListOf_append(&q->products, sref_clone((SpeciesReference_t *)r->products.items[j]));
    for (j = 0; j < r->modifiers.size; j++) ListOf_append(&q->modifiers, sref_clone((SpeciesReference_t *)r->modifiers.items[j]));
    if (r->kineticLaw) {
      KineticLaw_t *k = (KineticLaw_t *)calloc(1, sizeof *k);
      k->math = ASTNode_deepCopy(r->kineticLaw->math);
      for (j = 0; j < r->kineticLaw->parameters.size; j++) ListOf_append(&k->parameters, Parameter_clone((Parameter_t *)r->kineticLaw->parameters.items[j]));
      q->kineticLaw = k;
    }
  }
  for (i = 0; i < m->events.size; i++) ListOf_append(&n->events, Event_clone(Model_getEvent(m, i)));
  return n;
}

static void sref_free(SpeciesReference_t *s) { if (!s) return; free(s->species); ASTNode_free(s->stoichiometryMath); free(s); }
static void param_free(Parameter_t *p) { if (!p) return; free(p->id); free(p); }
void Model_free(Model_t *m)
{
  unsigned int i, j;
  if (!m) return;
  for (i = 0; i < m->functionDefinitions.size; i++) { FunctionDefinition_t *f = Model_getFunctionDefinition(m, i); free(f->id); ASTNode_free(f->math); free(f); }
  for (i = 0; i < m->compartments.size; i++) { Compartment_t *c = Model_getCompartment(m, i); free(c->id); free(c); }
  for (i = 0; i < m->species.size; i++) { Species_t *s = Model_getSpecies(m, i); free(s->id); free(s->compartment); free(s); }
  for (i = 0; i < m->parameters.size; i++) param_free(Model_getParameter(m, i));
  for (i = 0; i < m->initialAssignments.size; i++) { InitialAssignment_t *a = Model_getInitialAssignment(m, i); free(a->symbol); ASTNode_free(a->math); free(a); }
  for (i = 0; i < m->rules.size; i++) { Rule_t *r = Model_getRule(m, i); free(r->variable); ASTNode_free(r->math); free(r->formula); free(r); }
  for (i = 0; i < m->reactions.size; i++) {
    Reaction_t *r = Model_getReaction(m, i);
    for (j = 0; j < r->reactants.size; j++) sref_free((SpeciesReference_t *)r->reactants.items[j]);
    for (j = 0; j < r->products.size; j++) sref_free((SpeciesReference_t *)r->products.items[j]);
    for (j = 0; j < r->modifiers.size; j++) sref_free((SpeciesReference_t *)r->modifiers.items[j]);
    free(r->reactants.items); free(r->products.items); free(r->modifiers.items);
    if (r->kineticLaw) {
      ASTNode_free(r->kineticLaw->math);
      for (j = 0; j < r->kineticLaw->parameters.size; j++) param_free((Parameter_t *)r->kineticLaw->parameters.items[j]);
      free(r->kineticLaw->parameters.items); free(r->kineticLaw->formula); free(r->kineticLaw);
    }
    free(r->id); free(r);
  }
  for (i = 0; i < m->events.size; i++) {
    Event_t *e = Model_getEvent(m, i);
    for (j = 0; j < e->assignments.size; j++) { EventAssignment_t *a = (EventAssignment_t *)e->assignments.items[j]; free(a->variable); ASTNode_free(a->math); free(a); }
    free(e->assignments.items); free(e->id); ASTNode_free(e->trigger); ASTNode_free(e->delay); free(e);
  }
  free(m->functionDefinitions.items); free(m->compartments.items); free(m->species.items); free(m->parameters.items);
  free(m->initialAssignments.items); free(m->rules.items); free(m->reactions.items); free(m->events.items);
  free(m->id); free(m->name); free(m->notes); free(m);
}

SBMLDocument_t *SBMLDocument_create(void) { SBMLDocument_t *d = (SBMLDocument_t *)calloc(1, sizeof *d); d->level = 2; d->version = 1; return d; }
void SBMLDocument_free(SBMLDocument_t *d) { if (!d) return; Model_free(d->model); free(d); }
Model_t *SBMLDocument_getModel(SBMLDocument_t *d) { return d ? d->model : NULL; }
unsigned int SBMLDocument_getLevel(const SBMLDocument_t *d) { return d->level; }

/* ------------------------------------------------------ SBML -> model */
static const char *id_of(const xnode_t *x, unsigned int level)
{
  const char *id = (level == 1) ? xattr(x, "name") : xattr(x, "id");
  return id;
}
static ASTNode_t *math_of(const xnode_t *x, unsigned int level)
{
  if (level == 1) { const char *f = xattr(x, "formula"); return f ? SBML_parseFormula(f) : NULL; }
  return math_from_xnode(xchild(x, "math"));
}

static void read_parameters(const xnode_t *list, unsigned int level, ListOf_t *out)
{
  unsigned int i;
  if (!list) return;
  for (i = 0; i < list->nkids; i++) {
    const xnode_t *x = list->kids[i]; Parameter_t *p;
    if (strcmp(x->name, "parameter") != 0) continue;
    p = (Parameter_t *)calloc(1, sizeof *p);
    p->id = xstrdup(id_of(x, level));
    if (xattr(x, "value")) { p->value = xdouble(xattr(x, "value")); p->isSetValue = 1; }
    p->constant = xbool(xattr(x, "constant"), 1);
    ListOf_append(out, p);
  }
}
static void read_srefs(const xnode_t *list, unsigned int level, ListOf_t *out)
{
  unsigned int i;
  if (!list) return;
  for (i = 0; i < list->nkids; i++) {
    const xnode_t *x = list->kids[i]; SpeciesReference_t *s; const char *sp, *st; xnode_t *sm;
    if (strstr(x->name, "Reference") == NULL) continue;
    sp = xattr(x, "species"); if (!sp) sp = xattr(x, "specie");
    s = sref_new(sp, 1.0);
    st = xattr(x, "stoichiometry");
    if (st) s->stoichiometry = xdouble(st);
    if (level == 1 && xattr(x, "denominator")) s->stoichiometry /= xdouble(xattr(x, "denominator"));
    sm = xchild(x, "stoichiometryMath");
    if (sm) s->stoichiometryMath = math_from_xnode(xchild(sm, "math"));
    ListOf_append(out, s);
  }
}

static Model_t *model_from_xnode(const xnode_t *xm, unsigned int level, unsigned int version)
{
  unsigned int i;
  xnode_t *l;
  Model_t *m = Model_create(level >= 2 ? level : 2, level >= 2 ? version : 1);
  m->id = xstrdup(id_of(xm, level));
  m->name = xstrdup(xattr(xm, "name"));

  if ((l = xchild(xm, "listOfFunctionDefinitions")))
    for (i = 0; i < l->nkids; i++) {
      const xnode_t *x = l->kids[i];
      if (strcmp(x->name, "functionDefinition") != 0) continue;
      Model_createFunctionDefinition(m, xattr(x, "id"), math_from_xnode(xchild(x, "math")));
    }
  if ((l = xchild(xm, "listOfCompartments")))
    for (i = 0; i < l->nkids; i++) {
      const xnode_t *x = l->kids[i]; Compartment_t *c; const char *v;
      if (strcmp(x->name, "compartment") != 0) continue;
      c = (Compartment_t *)calloc(1, sizeof *c);
      c->id = xstrdup(id_of(x, level));
      c->spatialDimensions = xattr(x, "spatialDimensions") ? atoi(xattr(x, "spatialDimensions")) : 3;
      c->constant = xbool(xattr(x, "constant"), 1);
      v = (level == 1) ? xattr(x, "volume") : xattr(x, "size");
      if (v) { c->size = xdouble(v); c->isSetSize = 1; }
      else if (level == 1) { c->size = 1.0; c->isSetSize = 1; }   /* L1 default volume */
      ListOf_append(&m->compartments, c);
    }
  if ((l = xchild(xm, "listOfSpecies")))
    for (i = 0; i < l->nkids; i++) {
      const xnode_t *x = l->kids[i]; Species_t *s;
      if (strcmp(x->name, "species") != 0 && strcmp(x->name, "specie") != 0) continue;
      s = (Species_t *)calloc(1, sizeof *s);
      s->id = xstrdup(id_of(x, level));
      s->compartment = xstrdup(xattr(x, "compartment"));
      if (xattr(x, "initialAmount")) { s->initialAmount = xdouble(xattr(x, "initialAmount")); s->isSetInitialAmount = 1; }
      if (xattr(x, "initialConcentration")) { s->initialConcentration = xdouble(xattr(x, "initialConcentration")); s->isSetInitialConcentration = 1; }
      s->hasOnlySubstanceUnits = xbool(xattr(x, "hasOnlySubstanceUnits"), 0);
      s->boundaryCondition = xbool(xattr(x, "boundaryCondition"), 0);
      s->constant = xbool(xattr(x, "constant"), 0);
      ListOf_append(&m->species, s);
    }
  read_parameters(xchild(xm, "listOfParameters"), level, &m->parameters);
  if ((l = xchild(xm, "listOfInitialAssignments")))
    for (i = 0; i < l->nkids; i++) {
      const xnode_t *x = l->kids[i];
      if (strcmp(x->name, "initialAssignment") != 0) continue;
      Model_createInitialAssignment(m, xattr(x, "symbol"), math_from_xnode(xchild(x, "math")));
    }
  if ((l = xchild(xm, "listOfRules")))
    for (i = 0; i < l->nkids; i++) {
      const xnode_t *x = l->kids[i]; SBMLTypeCode_t type; const char *var = NULL;
      if (level == 1) {
        const char *t = xattr(x, "type");
        type = (t && strcmp(t, "rate") == 0) ? SBML_RATE_RULE : SBML_ASSIGNMENT_RULE;
        if (strcmp(x->name, "algebraicRule") == 0) type = SBML_ALGEBRAIC_RULE;
        else if (strcmp(x->name, "parameterRule") == 0) var = xattr(x, "name");
        else if (strcmp(x->name, "compartmentVolumeRule") == 0) var = xattr(x, "compartment");
        else { var = xattr(x, "species"); if (!var) var = xattr(x, "specie"); }
      } else {
        if (strcmp(x->name, "rateRule") == 0) type = SBML_RATE_RULE;
        else if (strcmp(x->name, "assignmentRule") == 0) type = SBML_ASSIGNMENT_RULE;
        else if (strcmp(x->name, "algebraicRule") == 0) type = SBML_ALGEBRAIC_RULE;
        else continue;
        var = xattr(x, "variable");
      }
      Model_createRule(m, type, var, math_of(x, level));
    }
  if ((l = xchild(xm, "listOfReactions")))
    for (i = 0; i < l->nkids; i++) {
      const xnode_t *x = l->kids[i]; Reaction_t *r; xnode_t *kl;
      if (strcmp(x->name, "reaction") != 0) continue;
      r = Model_createReaction(m, id_of(x, level));
      r->reversible = xbool(xattr(x, "reversible"), 1);
      r->fast = xbool(xattr(x, "fast"), 0);
      read_srefs(xchild(x, "listOfReactants"), level, &r->reactants);
      read_srefs(xchild(x, "listOfProducts"), level, &r->products);
      read_srefs(xchild(x, "listOfModifiers"), level, &r->modifiers);
      if ((kl = xchild(x, "kineticLaw"))) {
        KineticLaw_t *k = (KineticLaw_t *)calloc(1, sizeof *k);
        k->math = math_of(kl, level);
        read_parameters(xchild(kl, "listOfParameters"), level, &k->parameters);
        r->kineticLaw = k;
      }
    }
  if ((l = xchild(xm, "listOfEvents")))
    for (i = 0; i < l->nkids; i++) {
      const xnode_t *x = l->kids[i]; Event_t *e; xnode_t *t, *d, *la; unsigned int j;
      if (strcmp(x->name, "event") != 0) continue;
      t = xchild(x, "trigger"); d = xchild(x, "delay");
      e = Model_createEvent(m, xattr(x, "id"), t ? math_from_xnode(xchild(t, "math")) : NULL);
      if (d) e->delay = math_from_xnode(xchild(d, "math"));
      if ((la = xchild(x, "listOfEventAssignments")))
        for (j = 0; j < la->nkids; j++) {
          const xnode_t *a = la->kids[j];
          if (strcmp(a->name, "eventAssignment") != 0) continue;
          Event_createEventAssignment(e, xattr(a, "variable"), math_from_xnode(xchild(a, "math")));
        }
    }
  return m;
}

SBMLDocument_t *readSBMLFromString(const char *xml)
{
  xnode_t *root = xparse(xml), *xm;
  SBMLDocument_t *d;
  unsigned int level, version;
  if (!root) return NULL;
  if (strcmp(root->name, "sbml") != 0 || !(xm = xchild(root, "model"))) { xfree(root); return NULL; }
  level = xattr(root, "level") ? (unsigned int)atoi(xattr(root, "level")) : 2;
  version = xattr(root, "version") ? (unsigned int)atoi(xattr(root, "version")) : 1;
  d = SBMLDocument_create();
  d->level = level >= 2 ? level : 2;          /* L1 is converted on read, cf. src/sbml.c:109-118 */
  d->version = level >= 2 ? version : 1;
  d->model = model_from_xnode(xm, level, version);
  xfree(root);
  return d;
}

SBMLDocument_t *readSBML(const char *filename)
{
  FILE *f = fopen(filename, "rb");
  long n; char *buf; SBMLDocument_t *d;
  if (!f) return NULL;
  fseek(f, 0, SEEK_END); n = ftell(f); fseek(f, 0, SEEK_SET);
  buf = (char *)malloc((size_t)n + 1);
  if (fread(buf, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(buf); return NULL; }
  buf[n] = 0; fclose(f);
  d = readSBMLFromString(buf);
  free(buf);
  return d;
}

/* ------------------------------------------------------ model -> SBML */
static void write_math(wbuf_t *w, const ASTNode_t *n, const char *indent)
{
  char *s = writeMathMLToString(n);
  wprintf_(w, "%s", indent); wput(w, s); wput(w, "\n");
  free(s);
}
static void write_params(wbuf_t *w, const ListOf_t *l, const char *indent)
{
  unsigned int i;
  if (!l->size) return;
  wprintf_(w, "%s<listOfParameters>\n", indent);
  for (i = 0; i < l->size; i++) {
    const Parameter_t *p = (const Parameter_t *)l->items[i];
    wprintf_(w, "%s <parameter id=\"%s\"", indent, p->id);
    if (p->isSetValue) wprintf_(w, " value=\"%.17g\"", p->value);
    if (!p->constant) wput(w, " constant=\"false\"");
    wput(w, "/>\n");
  }
  wprintf_(w, "%s</listOfParameters>\n", indent);
}
static void write_srefs(wbuf_t *w, const ListOf_t *l, const char *tag, const char *item)
{
  unsigned int i;
  if (!l->size) return;
  wprintf_(w, "    <%s>\n", tag);
  for (i = 0; i < l->size; i++) {
    const SpeciesReference_t *s = (const SpeciesReference_t *)l->items[i];
    wprintf_(w, "     <%s species=\"%s\"", item, s->species);
    if (strcmp(item, "speciesReference") == 0 && !s->stoichiometryMath && s->stoichiometry != 1.0) wprintf_(w, " stoichiometry=\"%.17g\"", s->stoichiometry);
    if (s->stoichiometryMath) { wput(w, ">\n      <stoichiometryMath>\n"); write_math(w, s->stoichiometryMath, "       "); wprintf_(w, "      </stoichiometryMath>\n     </%s>\n", item); }
    else wput(w, "/>\n");
  }
  wprintf_(w, "    </%s>\n", tag);
}

char *writeSBMLToString(const SBMLDocument_t *d)
{
  wbuf_t w; const Model_t *m = d->model; unsigned int i, j;
  w.b = NULL; w.len = w.cap = 0;
  wput(&w, "<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n");
  wput(&w, "<sbml xmlns=\"http://www.sbml.org/sbml/level2\" level=\"2\" version=\"1\">\n");
  wprintf_(&w, " <model id=\"%s\">\n", m->id ? m->id : "model");
  if (m->functionDefinitions.size) {
    wput(&w, "  <listOfFunctionDefinitions>\n");
    for (i = 0; i < m->functionDefinitions.size; i++) { FunctionDefinition_t *f = Model_getFunctionDefinition(m, i); wprintf_(&w, "   <functionDefinition id=\"%s\">\n", f->id); write_math(&w, f->math, "    "); wput(&w, "   </functionDefinition>\n"); }
    wput(&w, "  </listOfFunctionDefinitions>\n");
  }
  if (m->compartments.size) {
    wput(&w, "  <listOfCompartments>\n");
    for (i = 0; i < m->compartments.size; i++) {
      Compartment_t *c = Model_getCompartment(m, i);
      wprintf_(&w, "   <compartment id=\"%s\"", c->id);
      if (c->isSetSize) wprintf_(&w, " size=\"%.17g\"", c->size);
      if (c->spatialDimensions != 3) wprintf_(&w, " spatialDimensions=\"%d\"", c->spatialDimensions);
      if (!c->constant) wput(&w, " constant=\"false\"");
      wput(&w, "/>\n");
    }
    wput(&w, "  </listOfCompartments>\n");
  }
  if (m->species.size) {
    wput(&w, "  <listOfSpecies>\n");
    for (i = 0; i < m->species.size; i++) {
      Species_t *s = Model_getSpecies(m, i);
      wprintf_(&w, "   <species id=\"%s\" compartment=\"%s\"", s->id, s->compartment ? s->compartment : "");
      if (s->isSetInitialAmount) wprintf_(&w, " initialAmount=\"%.17g\"", s->initialAmount);
      if (s->isSetInitialConcentration) wprintf_(&w, " initialConcentration=\"%.17g\"", s->initialConcentration);
      if (s->hasOnlySubstanceUnits) wput(&w, " hasOnlySubstanceUnits=\"true\"");
      if (s->boundaryCondition) wput(&w, " boundaryCondition=\"true\"");
      if (s->constant) wput(&w, " constant=\"true\"");
      wput(&w, "/>\n");
    }
    wput(&w, "  </listOfSpecies>\n");
  }
  if (m->parameters.size) write_params(&w, &m->parameters, "  ");
  if (m->initialAssignments.size) {
    wput(&w, "  <listOfInitialAssignments>\n");
    for (i = 0; i < m->initialAssignments.size; i++) { InitialAssignment_t *a = Model_getInitialAssignment(m, i); wprintf_(&w, "   <initialAssignment symbol=\"%s\">\n", a->symbol); write_math(&w, a->math, "    "); wput(&w, "   </initialAssignment>\n"); }
    wput(&w, "  </listOfInitialAssignments>\n");
  }
  if (m->rules.size) {
    wput(&w, "  <listOfRules>\n");
    for (i = 0; i < m->rules.size; i++) {
      Rule_t *r = Model_getRule(m, i);
      const char *tag = r->type == SBML_RATE_RULE ? "rateRule" : r->type == SBML_ASSIGNMENT_RULE ? "assignmentRule" : "algebraicRule";
      if (r->variable) wprintf_(&w, "   <%s variable=\"%s\">\n", tag, r->variable); else wprintf_(&w, "   <%s>\n", tag);
      write_math(&w, r->math, "    ");
      wprintf_(&w, "   </%s>\n", tag);
    }
    wput(&w, "  </listOfRules>\n");
  }
  if (m->reactions.size) {
    wput(&w, "  <listOfReactions>\n");
    for (i = 0; i < m->reactions.size; i++) {
      Reaction_t *r = Model_getReaction(m, i);
      wprintf_(&w, "   <reaction id=\"%s\"%s>\n", r->id, r->reversible ? "" : " reversible=\"false\"");
      write_srefs(&w, &r->reactants, "listOfReactants", "speciesReference");
      write_srefs(&w, &r->products, "listOfProducts", "speciesReference");
      write_srefs(&w, &r->modifiers, "listOfModifiers", "modifierSpeciesReference");
      if (r->kineticLaw) {
        wput(&w, "    <kineticLaw>\n");
        write_math(&w, r->kineticLaw->math, "     ");
        write_params(&w, &r->kineticLaw->parameters, "     ");
        wput(&w, "    </kineticLaw>\n");
      }
      wput(&w, "   </reaction>\n");
    }
    wput(&w, "  </listOfReactions>\n");
  }
  if (m->events.size) {
    wput(&w, "  <listOfEvents>\n");
    for (i = 0; i < m->events.size; i++) {
      Event_t *e = Model_getEvent(m, i);
      if (e->id) wprintf_(&w, "   <event id=\"%s\">\n", e->id); else wput(&w, "   <event>\n");
      wput(&w, "    <trigger>\n"); write_math(&w, e->trigger, "     "); wput(&w, "    </trigger>\n");
      if (e->delay) { wput(&w, "    <delay>\n"); write_math(&w, e->delay, "     "); wput(&w, "    </delay>\n"); }
      wput(&w, "    <listOfEventAssignments>\n");
      for (j = 0; j < e->assignments.size; j++) {
        EventAssignment_t *a = (EventAssignment_t *)e->assignments.items[j];
        wprintf_(&w, "     <eventAssignment variable=\"%s\">\n", a->variable); write_math(&w, a->math, "      "); wput(&w, "     </eventAssignment>\n");
      }
      wput(&w, "    </listOfEventAssignments>\n   </event>\n");
    }
    wput(&w, "  </listOfEvents>\n");
  }
  wput(&w, " </model>\n</sbml>\n");
  return w.b;
}

int writeSBML(const SBMLDocument_t *d, const char *filename)
{
  char *s = writeSBMLToString(d);
  FILE *f = fopen(filename, "wb");
  if (!f) { free(s); return 0; }
  fputs(s, f); fclose(f); free(s);
  return 1;
}

/* libSBML's reader object (callers create one, read, free it) */
SBMLReader_t *SBMLReader_create(void) { return (SBMLReader_t *)calloc(1, sizeof(SBMLReader_t)); }
SBMLDocument_t *SBMLReader_readSBML(SBMLReader_t *sr, const char *filename) { (void)sr; return readSBML(filename); }
SBMLDocument_t *SBMLReader_readSBMLFromString(SBMLReader_t *sr, const char *xml) { (void)sr; return readSBMLFromString(xml); }
void SBMLReader_free(SBMLReader_t *sr) { free(sr); }

/* ------------------------------------------------------ libSBML-named accessors */
const char *Compartment_getId(const Compartment_t *c) { return c->id; }
const char *Species_getId(const Species_t *s) { return s->id; }
const char *Parameter_getId(const Parameter_t *p) { return p->id; }
double Parameter_getValue(const Parameter_t *p) { return p->value; }
void Parameter_setValue(Parameter_t *p, double v) { p->value = v; p->isSetValue = 1; }
void Parameter_setName(Parameter_t *p, const char *name) { (void)p; (void)name; }       /* names are not kept */
void Parameter_setConstant(Parameter_t *p, int c) { p->constant = c; }
void Compartment_setConstant(Compartment_t *c, int v) { c->constant = v; }
Parameter_t *Parameter_create(void) { Parameter_t *p = (Parameter_t *)calloc(1, sizeof *p); p->constant = 1; return p; }
Parameter_t *Parameter_createWith(const char *id, const char *name) { Parameter_t *p = Parameter_create(); (void)name; Parameter_setId(p, id); return p; }
void Parameter_free(Parameter_t *p) { param_free(p); }
const char *Reaction_getId(const Reaction_t *r) { return r->id; }
KineticLaw_t *Reaction_getKineticLaw(const Reaction_t *r) { return r->kineticLaw; }
const ASTNode_t *KineticLaw_getMath(const KineticLaw_t *k) { return k->math; }
const char *KineticLaw_getFormula(const KineticLaw_t *ck)
{
  KineticLaw_t *k = (KineticLaw_t *)ck;
  free(k->formula);
  k->formula = k->math ? SBML_formulaToString(k->math) : NULL;
  return k->formula;
}
const char *Rule_getFormula(const Rule_t *cr)
{
  Rule_t *r = (Rule_t *)cr;
  free(r->formula);
  r->formula = r->math ? SBML_formulaToString(r->math) : NULL;
  return r->formula;
}
const char *Rule_getVariable(const Rule_t *r) { return r->variable; }
unsigned int Model_getNumUnitDefinitions(const Model_t *m) { (void)m; return 0; }
unsigned int Model_getNumSpeciesTypes(const Model_t *m) { (void)m; return 0; }
unsigned int Model_getNumCompartmentTypes(const Model_t *m) { (void)m; return 0; }
unsigned int Model_getNumConstraints(const Model_t *m) { (void)m; return 0; }
unsigned int Model_getNumSpeciesWithBoundaryCondition(const Model_t *m)
{
  unsigned int i, n = 0;
  for (i = 0; i < m->species.size; i++) if (Model_getSpecies(m, i)->boundaryCondition) n++;
  return n;
}
ListOf_t *ListOf_create(void) { return (ListOf_t *)calloc(1, sizeof(ListOf_t)); }
void ListOf_appendAndOwn(ListOf_t *l, SBase_t *item) { ListOf_append(l, item); }
void ListOf_free(ListOf_t *l) { if (l) { free(l->items); free(l); } }
