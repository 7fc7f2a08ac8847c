import { register } from 'node:module';
register('./ts_resolver.mjs', import.meta.url);
