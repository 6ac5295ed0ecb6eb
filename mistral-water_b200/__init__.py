"""mistral-water_b200: B200-native engine behind the Mistral Water asset's ocean / pond hot path.

The product is libmistral_ocean.so (hand-written sm_100a CUDA behind a C ABI, include/mistral_ocean.h).
This package is the host-side mirror of the reference's C# surface, bound with ctypes:

    FFTMesh        -- Scripts/FFTMesh.cs MonoBehaviour (fields, Awake/Update/EvaluateWaves)
    Ocean          -- one mw_ocean handle (host or device buffers, batched tiles)
    GerstnerWaves  -- Shaders/MistralWaterLib.cginc Gerstner / GerstnerLevelOne
    OceanRenderer  -- Scripts/OceanRenderer.cs MonoBehaviour (the GPU-shader convention: four R x R maps per frame)
    tiles          -- one-tile-per-GPU sharding + all-gather over torch.distributed (NCCL)

Importing it loads the shared library and fails loudly if it is not built: there is no CPU path.
"""
from . import native

native.load()

from .ocean import Ocean, fft2d  # noqa: E402
from .fft_mesh import FFTMesh, Mesh  # noqa: E402
from .pond import GerstnerWaves, POND_MATERIAL, pond_wave_table_32, wave_displace  # noqa: E402
from .ocean_renderer import OceanRenderer, Renderer, generate_mesh  # noqa: E402

__all__ = ["native", "Ocean", "fft2d", "FFTMesh", "Mesh", "GerstnerWaves", "POND_MATERIAL", "pond_wave_table_32",
           "OceanRenderer", "Renderer", "generate_mesh", "wave_displace"]
