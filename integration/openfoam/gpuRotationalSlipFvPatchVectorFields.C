#include "gpuRotationalSlipFvPatchVectorFields.H"
#include "addToRunTimeSelectionTable.H"

namespace Foam
{
    makePatchTypeField(fvPatchVectorField, constantRotationalSlipFvPatchVectorField);
    makePatchTypeField(fvPatchVectorField, exponentialRotationalSlipFvPatchVectorField);
}
