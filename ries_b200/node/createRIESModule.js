// Drop-in for the emscripten factory used by src/workers/ries.worker.js:64-123: same object shape, the search
// runs on the GPU through libries_b200.so.  Output lines reach `print` exactly as the wasm module delivers them.
const addon = require('./build/Release/ries_b200.node');

async function createRIESModule({ print = console.log } = {}) {
  return {
    callMain(args) {
      const r = addon.callMain(args.map(String));
      const lines = r.stdout.split('\n');
      if (lines.length && lines[lines.length - 1] === '') lines.pop();
      lines.forEach((l) => print(l));
      return r.code;
    },
  };
}

module.exports = createRIESModule;
