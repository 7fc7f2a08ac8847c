// Module-resolution hook: the reference imports its own files without extensions (e.g. src/index.ts:20-37
// `from './binaryQuantizationFormat'`); map './x' -> './x.ts' when that file exists.
import { existsSync } from 'node:fs';
import { fileURLToPath, pathToFileURL } from 'node:url';
import { dirname, resolve as pathResolve } from 'node:path';

export async function resolve(specifier, context, nextResolve) {
  if ((specifier.startsWith('./') || specifier.startsWith('../')) && context.parentURL && !/\.[cm]?[jt]s$/.test(specifier)) {
    const base = pathResolve(dirname(fileURLToPath(context.parentURL)), specifier);
    for (const cand of [base + '.ts', pathResolve(base, 'index.ts')])
      if (existsSync(cand)) return { url: pathToFileURL(cand).href, shortCircuit: true, format: 'module-typescript' };
  }
  return nextResolve(specifier, context);
}
