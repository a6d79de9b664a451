/*
 * sbmlResults.h -- simulation results mapped back onto SBML structures: time courses of all species, of the
 * non-constant compartments and global parameters, and of the reaction fluxes.
 * Reference: src/sbmlsolver/sbmlResults.h:66-135 (same struct layouts and entry points), src/sbmlResults.c:66-560,
 * src/odeSolver.c:441-543 (SBMLResults_fromIntegrator).  Forward sensitivities are mapped by SBMLResults_createSens' restatement
 * (src/odeSolver.c:545-575).
 */
#ifndef ODES_B200_SBMLRESULTS_H
#define ODES_B200_SBMLRESULTS_H
#include "sbmlsolver/sbmlModel.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct timeCourse timeCourse_t;
typedef struct timeCourseArray timeCourseArray_t;
typedef struct _SBMLResults SBMLResults_t;
typedef struct _SBMLResultsMatrix SBMLResultsMatrix_t;
typedef struct _SBMLResultsArray SBMLResultsArray_t;

struct timeCourse {
  int timepoints;        /* number of time points, including the initial one */
  char *name;            /* SBML id */
  double *values;        /* [timepoints] */
  double **sensitivity;  /* [nsens][timepoints] of an ODE variable when sensitivities were computed, else NULL */
};
struct timeCourseArray { int num_val; timeCourse_t **tc; };
struct _SBMLResults {
  timeCourse_t *time;
  timeCourseArray_t *species;        /* all species */
  timeCourseArray_t *compartments;   /* non-constant compartments */
  timeCourseArray_t *parameters;     /* non-constant global parameters */
  timeCourseArray_t *fluxes;         /* one per reaction */
  int nsens; char **param;           /* ids of the sensitivity parameters (src/odeSolver.c:545-575) */
};
struct _SBMLResultsMatrix { SBMLResults_t ***results; int i; int j; };
struct _SBMLResultsArray { SBMLResults_t **results; int size; };

timeCourse_t *SBMLResults_getTime(SBMLResults_t *);
timeCourse_t *SBMLResults_getTimeCourse(SBMLResults_t *, const char *);
int SBMLResults_getNout(const SBMLResults_t *);
int SBMLResults_getNumSens(const SBMLResults_t *);
const char *SBMLResults_getSensParam(const SBMLResults_t *, int);
timeCourse_t *Compartment_getTimeCourse(Compartment_t *, SBMLResults_t *);
timeCourse_t *Species_getTimeCourse(Species_t *, SBMLResults_t *);
timeCourse_t *Parameter_getTimeCourse(Parameter_t *, SBMLResults_t *);
const char *TimeCourse_getName(const timeCourse_t *);
int TimeCourse_getNumValues(const timeCourse_t *);
double TimeCourse_getValue(const timeCourse_t *, int);
double TimeCourse_getSensitivity(const timeCourse_t *, int, int);
void SBMLResults_dump(const SBMLResults_t *);
/* reference sbmlResults.h:134 (declared there, never defined): writes the LAST stored value of every species, non-constant
   compartment and non-constant global parameter course into the model's initial values, so that a following run continues
   from the end state; returns the model */
Model_t *Model_setValues(Model_t *, SBMLResults_t *);
void SBMLResults_dumpSpecies(const SBMLResults_t *);
void SBMLResults_dumpCompartments(const SBMLResults_t *);
void SBMLResults_dumpParameters(const SBMLResults_t *);
void SBMLResults_dumpFluxes(const SBMLResults_t *);
void SBMLResults_free(SBMLResults_t *);
void SBMLResultsArray_free(SBMLResultsArray_t *);
int SBMLResultsArray_getNumResults(const SBMLResultsArray_t *);
SBMLResults_t *SBMLResultsArray_getResults(SBMLResultsArray_t *, int i);
SBMLResults_t *SBMLResultsMatrix_getResults(SBMLResultsMatrix_t *, int i, int j);

/* construction (non-API in the reference, src/sbmlResults.c:322-430) */
SBMLResultsArray_t *SBMLResultsArray_allocate(int size);
SBMLResults_t *SBMLResults_create(Model_t *, int timepoints);

#ifdef __cplusplus
}
#endif
#endif
