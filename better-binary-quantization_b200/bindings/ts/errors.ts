/**
 * Maps the addon's error code ("BBQ:<status>:<vector>:<position>", see bindings/napi/bbq_napi.c) back to the
 * exact `throw new Error(...)` text of the reference, so callers that match on messages keep working.
 * Status values: include/bbq_b200.h (bbq_status).
 */
export function toReferenceError(e: any, what: 'build' | 'search'): Error {
  const m = /^BBQ:(\d+):(-?\d+):(-?\d+)$/.exec(e?.code ?? '');
  if (!m) return e;
  const [status, vec, pos] = [Number(m[1]), Number(m[2]), Number(m[3])];
  switch (status) {
    case 1: return new Error('queryBits必须在1-8之间');            // binaryQuantizationFormat.ts:144
    case 2: return new Error('indexBits必须在1-8之间');            // :147
    case 3: return new Error('向量集合不能为空');                    // :170
    case 4: return new Error('查询向量维度与目标向量维度不匹配');      // :328
    case 5: return new Error(what === 'build' ? `向量 ${vec} 位置 ${pos} 包含NaN值`       // :203
                                              : `向量位置 ${pos} 包含NaN值`);             // optimizedScalarQuantizer.ts:142
    case 6: return new Error(what === 'build' ? `向量 ${vec} 位置 ${pos} 包含Infinity值`  // :206
                                              : `向量位置 ${pos} 包含Infinity值`);        // optimizedScalarQuantizer.ts:145
    case 7: return new Error('k值不能为负数');                      // :325
    case 8: return new Error(what === 'search' ? '查询向量不能为空' : '目标向量集合不能为空'); // :319,:322
    default: return new Error(`[bbq-b200] ${e.message} (status ${status})`);              // CUDA / unsupported: no fallback
  }
}
