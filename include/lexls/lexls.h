// lexls/lexls.h — the reference's thin public wrapper (include/lexls/lexls.h:16-68): LexLS::LexLSE with resize/setObjDim/solve.
#pragma once

#include <lexls/lexlse.h>
#include <lexls/lexlsi.h>

namespace LexLS
{
class LexLSE
{
public:
    LexLSE() {}
    void resize(Index nVar, Index nObj, Index *maxObjDim) { impl.resize(nVar, nObj, maxObjDim); }
    void setObjDim(Index *ObjDim) { impl.setObjDim(ObjDim); }
    void setData(Index ObjIndex, const dMatrixConstRef &data) { impl.setData(ObjIndex, data); } // addition: the reference wrapper has none
    const dVectorType &solve()
    {
        impl.factorize();
        impl.solve();
        return impl.get_x();
    }
    internal::LexLSE impl;
};
} // namespace LexLS
