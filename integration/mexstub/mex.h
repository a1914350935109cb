/*
 * mex.h -- declarations of the part of the MATLAB / Octave MEX C API that the reference's mcxlabcl.cpp uses, so that the
 * UNCHANGED file can be compiled and linked against the B200 binding in an image that has neither MATLAB nor Octave
 * (integration/build_cli.py::build_mexcheck).  Declarations only: the resulting object leaves every mx / mex symbol
 * undefined, exactly as a real MEX file does until MATLAB loads it.  Names and signatures are those of the public
 * MEX API (matrix.h / mex.h of MATLAB R2018a+ with the interleaved-complex accessors mcxlabcl.cpp calls).
 */
#ifndef MCXB_MEXSTUB_H
#define MCXB_MEXSTUB_H

#include <stddef.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef char mxChar;
typedef bool mxLogical;

typedef enum { mxUNKNOWN_CLASS = 0, mxCELL_CLASS, mxSTRUCT_CLASS, mxLOGICAL_CLASS, mxCHAR_CLASS, mxVOID_CLASS, mxDOUBLE_CLASS, mxSINGLE_CLASS,
               mxINT8_CLASS, mxUINT8_CLASS, mxINT16_CLASS, mxUINT16_CLASS, mxINT32_CLASS, mxUINT32_CLASS, mxINT64_CLASS, mxUINT64_CLASS,
               mxFUNCTION_CLASS
             } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX } mxComplexity;
typedef struct { float real, imag; } mxComplexSingle;

void mexErrMsgTxt(const char* msg);
void mexWarnMsgTxt(const char* msg);
int mexPrintf(const char* fmt, ...);
int mexEvalString(const char* cmd);

double* mxGetPr(const mxArray* a);
void* mxGetData(const mxArray* a);
void* mxGetImagData(const mxArray* a);
mxComplexSingle* mxGetComplexSingles(const mxArray* a);
size_t mxGetNumberOfElements(const mxArray* a);
const mwSize* mxGetDimensions(const mxArray* a);
mwSize mxGetNumberOfDimensions(const mxArray* a);
int mxGetString(const mxArray* a, char* buf, mwSize buflen);
bool mxIsChar(const mxArray* a);
bool mxIsStruct(const mxArray* a);
bool mxIsDouble(const mxArray* a);
bool mxIsSingle(const mxArray* a);
bool mxIsUint8(const mxArray* a);
bool mxIsInt8(const mxArray* a);
bool mxIsUint16(const mxArray* a);
bool mxIsInt16(const mxArray* a);
bool mxIsUint32(const mxArray* a);
bool mxIsInt32(const mxArray* a);
bool mxIsUint64(const mxArray* a);
bool mxIsInt64(const mxArray* a);

int mxGetNumberOfFields(const mxArray* a);
int mxGetFieldNumber(const mxArray* a, const char* name);
const char* mxGetFieldNameByNumber(const mxArray* a, int n);
mxArray* mxGetFieldByNumber(const mxArray* a, mwIndex i, int n);
void mxSetFieldByNumber(mxArray* a, mwIndex i, int n, mxArray* v);
void mxSetField(mxArray* a, mwIndex i, const char* name, mxArray* v);
int mxAddField(mxArray* a, const char* name);

mxArray* mxCreateNumericArray(mwSize ndim, const mwSize* dims, mxClassID cls, mxComplexity cplx);
mxArray* mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char** names);
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity cplx);
mxArray* mxCreateDoubleScalar(double v);
mxArray* mxCreateString(const char* s);

/* the entry point a MEX file exports */
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);

#ifdef __cplusplus
}
#endif
#endif
