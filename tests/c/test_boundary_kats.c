/* Boundary checks a drop-in caller depends on (round 2 review): struct cvodeSettings has the reference's full layout
   (src/sbmlsolver/integratorSettings.h:56-148), the callback-seam entries of odeModel.h:356-363 exist and fail the way
   the reference's do, DetectNegState is refused instead of ignored, determinantNAST (processAST.h:62) and
   Model_setValues (sbmlResults.h:134) are there.  argv[1]: MAPK model file; argv[2] "cpu": no device on this box. */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "sbmlsolver/odeSolver.h"
#include "sbmlsolver/processAST.h"
#include "sbmlsolver/solverError.h"
#include "sbmlsolver/batchIntegrator.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

/* the reference's struct, restated field by field from its header, to compare sizes and offsets with */
struct ref_settings {
  double Time; int PrintStep; int StepResolution; double *TimePoints; int Indefinitely; double Error; double RError; int Mxstep;
  int DetectNegState; int CvodeMethod; int IterMethod; int MaxOrder; int ResetCvodeOnEvent; int SetTStop; int Sensitivity;
  char **sensIDs; int nsens; int SensMethod; int HaltOnEvent; int SteadyState; double ssThreshold; int UseJacobian; int StoreResults;
  int compileFunctions; int observation_data_type; int DoAdjoint; double AdjTime; int AdjPrintStep; double *AdjTimePoints;
  int nSaveSteps; int ncheck; double AdjError; double AdjRError; int AdjStoreResults; int OffSet; int InterStep; int doFIM;
};

int main(int argc, char **argv)
{
  const int cpu = argc > 2 && strcmp(argv[2], "cpu") == 0;
  CHECK(argc > 1);
  CHECK(sizeof(struct cvodeSettings) == sizeof(struct ref_settings));
  CHECK(offsetof(struct cvodeSettings, DoAdjoint) == offsetof(struct ref_settings, DoAdjoint));
  CHECK(offsetof(struct cvodeSettings, AdjTimePoints) == offsetof(struct ref_settings, AdjTimePoints));
  CHECK(offsetof(struct cvodeSettings, doFIM) == offsetof(struct ref_settings, doFIM));
  CHECK(SOLVER_MESSAGE_STEADYSTATE_FOUND == 120501 && SOLVER_MESSAGE_RERUN_WITH_OR_WO_JACOBIAN == 120500);

  odeModel_t *om = ODEModel_createFromFile(argv[1]);
  CHECK(om != NULL);

  /* callback seam: the Jacobian getter before the Jacobian exists fails with the reference's code */
  SolverError_clear();
  CHECK(ODEModel_getCompiledCVODEJacobianFunction(om) == NULL);
  CHECK(SolverError_getCode(ERROR_ERROR_TYPE, 0) == SOLVER_ERROR_CANNOT_COMPILE_JACOBIAN_NOT_COMPUTED);
  SolverError_clear();

  /* DetectNegState: refused at plan creation, with an error on the stack */
  {
    cvodeSettings_t *set = CvodeSettings_createWithTime(100.0, 10);
    int vary0 = 0;
    CvodeSettings_setDetectNegState(set, 1);
    SolverError_clear();
    CHECK(ODESBatch_create(om, set, 1, &vary0, 0, NULL, 0) == NULL);
    CHECK(SolverError_getNum(ERROR_ERROR_TYPE) == 1);
    CHECK(strstr(SolverError_getMessage(ERROR_ERROR_TYPE, 0), "DetectNegState") != NULL);
    SolverError_clear();
    CvodeSettings_setDetectNegState(set, 0);
    odesBatch_t *b = ODESBatch_create(om, set, 1, &vary0, 0, NULL, 0);
    CHECK(b != NULL);
    ODESBatch_free(b);
    CvodeSettings_free(set);
  }

  /* callback seam (plan creation above has constructed the Jacobian) */
  CHECK(ODEModel_constructJacobian(om) == 1);
  CHECK(ODEModel_compileCVODEFunctions(om) == 1);             /* generates + compiles for sm_100a; needs no device */
  {
    void *rhs = ODEModel_getCompiledCVODERHSFunction(om), *jac = ODEModel_getCompiledCVODEJacobianFunction(om);
    if (cpu) { CHECK(rhs == NULL && jac == NULL); CHECK(SolverError_getNum(ERROR_ERROR_TYPE) > 0); SolverError_clear(); }
    else {
      long long c0[8], c1[8];
      CHECK(rhs != NULL && jac == rhs);
      ODES_getCounters(c0);
      CHECK(ODEModel_getCompiledCVODERHSFunction(om) == rhs);   /* cached on the model: nothing compiled or loaded again */
      ODES_getCounters(c1);
      CHECK(c1[0] == c0[0] && c1[2] == c0[2]);
    }
  }
  {
    cvodeSettings_t *set = CvodeSettings_create();
    char *ids[2] = {"K1", "Ki"};
    odeSense_t *os;
    CvodeSettings_setSensitivity(set, 1);
    CvodeSettings_setSensParams(set, ids, 2);
    os = ODESense_create(om, set);
    CHECK(os != NULL);
    CHECK(ODESense_compileCVODESenseFunctions(os) == 1);
    if (!cpu) CHECK(ODESense_getCompiledCVODESenseFunction(os) != NULL);
    SolverError_clear();
    ODESense_free(os);
    CvodeSettings_free(set);
  }

  /* determinantNAST on a 2 x 2 and a 3 x 3 matrix of expressions, evaluated */
  {
    const char *e2[2][2] = {{"a", "b"}, {"c", "d"}};
    const char *e3[3][3] = {{"a", "0", "b"}, {"c", "d", "0"}, {"0", "a + 1", "c"}};
    char *names[4] = {"a", "b", "c", "d"};
    double vals[4] = {2.0, 3.0, 5.0, 7.0};
    ASTNode_t **rows2[2], *m2[2][2], **rows3[3], *m3[3][3], *det;
    int i, j;
    for (i = 0; i < 2; i++) { rows2[i] = m2[i]; for (j = 0; j < 2; j++) m2[i][j] = SBML_parseFormula(e2[i][j]); }
    for (i = 0; i < 3; i++) { rows3[i] = m3[i]; for (j = 0; j < 3; j++) m3[i][j] = SBML_parseFormula(e3[i][j]); }
    det = determinantNAST(rows2, 2);
    CHECK(det != NULL);
    { ASTNode_t *ix = indexAST(det, 4, names); cvodeData_t d; memset(&d, 0, sizeof d); d.value = vals; d.nvalues = 4;
      CHECK(fabs(evaluateAST(ix, &d) - (2.0 * 7.0 - 3.0 * 5.0)) < 1e-12); ASTNode_free(ix); }
    ASTNode_free(det);
    det = determinantNAST(rows3, 3);
    CHECK(det != NULL);
    /* a (d c - 0) - 0 + b (c (a + 1) - 0) = 2*35 + 3*15 = 115 */
    { ASTNode_t *ix = indexAST(det, 4, names); cvodeData_t d; memset(&d, 0, sizeof d); d.value = vals; d.nvalues = 4;
      CHECK(fabs(evaluateAST(ix, &d) - 115.0) < 1e-12); ASTNode_free(ix); }
    ASTNode_free(det);
    for (i = 0; i < 2; i++) for (j = 0; j < 2; j++) ASTNode_free(m2[i][j]);
    for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) ASTNode_free(m3[i][j]);
    CHECK(determinantNAST(NULL, 0) == NULL);
  }

  /* Model_setValues: the end of a stored course becomes the model's initial value */
  {
    SBMLReader_t *rd = SBMLReader_create();
    SBMLDocument_t *doc = SBMLReader_readSBML(rd, argv[1]);
    Model_t *m = SBMLDocument_getModel(doc);
    SBMLResults_t *r = SBMLResults_create(m, 3);
    timeCourse_t *tc;
    CHECK(m != NULL && r != NULL);
    tc = SBMLResults_getTimeCourse(r, "MAPK_PP");
    CHECK(tc != NULL);
    tc->values[0] = 1.0; tc->values[1] = 2.0; tc->values[2] = 42.5;
    CHECK(Model_setValues(m, r) == m);
    { const Species_t *sp = Model_getSpeciesById(m, "MAPK_PP");      /* the model gives initial amounts */
      CHECK((sp->isSetInitialAmount && !sp->isSetInitialConcentration ? sp->initialAmount : sp->initialConcentration) == 42.5); }
    SBMLResults_free(r);
    SBMLDocument_free(doc);
    SBMLReader_free(rd);
  }
  ODEModel_free(om);
  printf("ok\n");
  return 0;
}
