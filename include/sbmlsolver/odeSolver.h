/*
 * odeSolver.h -- high-level batch interface: varySettings_t and
 * Model_odeSolverBatch.  Reference: src/sbmlsolver/odeSolver.h:53-92,
 * src/odeSolver.c:235-352 (batch loop), :354-440 (globalize/localize).
 * Model_odeSolverBatch keeps its signature and meaning (one SBMLResults_t per
 * design point) but runs all design points in one batched device launch;
 * Model_odeSolverBatchFlat returns the same run as one flat batchResults_t
 * without the per-point containers (the form to use for 10^5..10^6 points).
 */
#ifndef ODES_B200_ODESOLVER_H
#define ODES_B200_ODESOLVER_H
#include "sbmlsolver/batchIntegrator.h"
#include "sbmlsolver/sbmlResults.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct varySettings varySettings_t;
struct varySettings {
  int nrdesignpoints; int nrparams; char **id; char **rid; double **params;
  int cnt_params; int cnt_points;
};

/* results of a batch run: all model values at all output times for every design point */
typedef struct batchResults {
  int nsets, nout, nvalues;      /* nout = number of time points (PrintStep+1) */
  char **names;                  /* [nvalues] */
  double *time;                  /* [nout] */
  double *value;                 /* [nsets][nout][nvalues] */
  int *status;                   /* [nsets] */
  /* forward sensitivities when set->Sensitivity was on: sens[set][time][parameter][equation], else nsens == 0 */
  int nsens, neq; char **sensParam; double *sens;
  /* reaction fluxes evaluated on the device from the stored rows: flux[set][time][reaction], names = reaction ids
     (nflux == 0: not computed) */
  int nflux; char **fluxNames; double *flux;
} batchResults_t;

varySettings_t *VarySettings_allocate(int nrparams, int nrdesignpoints);
int VarySettings_addDesignPoint(varySettings_t *, const double *);
int VarySettings_addParameter(varySettings_t *, const char *, const char *);
int VarySettings_setName(varySettings_t *, int, const char *, const char *);
int VarySettings_setValue(varySettings_t *, int, int, double);
double VarySettings_getValue(varySettings_t *, int, int);
int VarySettings_setValueByID(varySettings_t *, int, const char *, const char *, double);
double VarySettings_getValueByID(varySettings_t *, int, const char *, const char *);
const char *VarySettings_getName(const varySettings_t *, int);
const char *VarySettings_getReactionName(const varySettings_t *, int);
void VarySettings_dump(varySettings_t *);
void VarySettings_free(varySettings_t *);

SBMLResults_t *SBML_odeSolver(SBMLDocument_t *, cvodeSettings_t *);
SBMLResults_t *Model_odeSolver(Model_t *, cvodeSettings_t *);
SBMLResultsArray_t *SBML_odeSolverBatch(SBMLDocument_t *, cvodeSettings_t *, varySettings_t *);
SBMLResultsArray_t *Model_odeSolverBatch(Model_t *, cvodeSettings_t *, varySettings_t *);
SBMLResults_t *SBMLResults_fromIntegrator(Model_t *, integratorInstance_t *);
batchResults_t *Model_odeSolverBatchFlat(Model_t *, cvodeSettings_t *, varySettings_t *);
batchResults_t *SBML_odeSolverBatchFlat(SBMLDocument_t *, cvodeSettings_t *, varySettings_t *);
SBMLResultsArray_t *SBMLResultsArray_fromBatch(Model_t *, odeModel_t *, const batchResults_t *);
double BatchResults_getValue(const batchResults_t *, int set, const char *name, int timestep);
double BatchResults_getFlux(const batchResults_t *, int set, const char *reaction, int timestep);
/* the indexed kinetic laws the flux courses are evaluated from (src/odeSolver.c:463-471), one per reaction or NULL */
int SBMLResults_fluxASTs(Model_t *m, const odeModel_t *om, ASTNode_t ***kls);
void BatchResults_free(batchResults_t *);
int Model_globalizeParameter(Model_t *m, const char *id, const char *rid);
int Model_localizeParameter(Model_t *m, const char *id, const char *rid);

#ifdef __cplusplus
}
#endif
#endif
