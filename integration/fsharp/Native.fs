// Native.fs — P/Invoke surface of libps_b200.so (include/ps_b200.h) for PhysicsSimulatorCore.
//
// Add to PhysicsSimulatorCore.fsproj after `JointsRestorer.fs` and before `NativeBatch.fs` / `Simulator.fs`:
//     <Compile Include="Native.fs" />
//     <Compile Include="NativeBatch.fs" />
// NOT COMPILED in the image this repository is built in (no .NET SDK there): it is the binding a maintainer adds,
// written against the reference's own types; the same calls are exercised by physicssimulator_b200/_native.py.
namespace PhysicsSimulator

open System
open System.Collections.Generic
open System.Runtime.InteropServices

/// ps_step_config (include/ps_b200.h): StepConfiguration (Configuration.fs:28-55) + SimulationStepInterval, flattened.
[<Struct; StructLayout(LayoutKind.Sequential)>]
type PsStepConfig =
    val mutable epsilon: float
    val mutable baumgarte: float
    val mutable allowedPenetration: float
    val mutable maxCorrectionVelocity: float
    val mutable dt: float
    val mutable solverIterations: int
    val mutable enableFriction: int
    val mutable integratorKind: int // 0 FirstOrder, 1 AugmentedSecondOrder (Configuration.fs:5-7)
    val mutable broadphaseKind: int // 0 Dummy, 1 SpatialTree (Configuration.fs:24-26)
    val mutable restoreJoints: int // 0: Simulator.fs:55 stays commented out
    val mutable recordContacts: int // 1 when Collisions are read back (Gui/Scenes/SceneHost.fs:16-27)
    val mutable exactAllPairs: int
    val mutable treeLeafCapacity: int
    val mutable treeMaxDepth: int
    val mutable aabb_mode: int
    val mutable treeMinX: float
    val mutable treeMinY: float
    val mutable treeMinZ: float
    val mutable treeMaxX: float
    val mutable treeMaxY: float
    val mutable treeMaxZ: float

/// ps_bodies_view: caller-owned SoA arrays, borrowed for the duration of the call.
[<Struct; StructLayout(LayoutKind.Sequential)>]
type PsBodiesView =
    val mutable nWorlds: int
    val mutable worldOffset: nativeint // int32[nWorlds+1]
    val mutable shape: nativeint // int32[n]: 0 sphere, 1 box
    val mutable dims: nativeint // float[3n]
    val mutable mass: nativeint // float[n], +infinity = Mass.Infinite
    val mutable inertiaInv: nativeint // float[3n]: diagonal of PrincipalRotationalInertiaInverse
    val mutable pos: nativeint
    val mutable vel: nativeint
    val mutable rot: nativeint // float[9n], ROW-major
    val mutable angMom: nativeint
    val mutable force: nativeint
    val mutable torque: nativeint
    val mutable elasticity: nativeint
    val mutable muKinetic: nativeint
    val mutable muStatic: nativeint

module internal Native =
    [<Literal>]
    let Lib = "ps_b200" // libps_b200.so on the loader path

    [<DllImport(Lib)>]
    extern int ps_init(int device)

    [<DllImport(Lib)>]
    extern nativeint ps_last_error()

    [<DllImport(Lib)>]
    extern int ps_batch_create(nativeint& handle, PsBodiesView& bodies, nativeint joints, PsStepConfig& config)

    [<DllImport(Lib)>]
    extern int ps_batch_destroy(nativeint handle)

    [<DllImport(Lib)>]
    extern int ps_batch_step(nativeint handle, int nSteps)

    [<DllImport(Lib)>]
    extern int ps_batch_step_host(nativeint handle, int nSteps, float[] posIn, float[] velIn, float[] rotIn, float[] angMomIn, float[] posOut, float[] velOut, float[] rotOut, float[] angMomOut)

    [<DllImport(Lib)>]
    extern int ps_batch_get_state(nativeint handle, int w0, int w1, float[] pos, float[] vel, float[] rot, float[] angMom)

    [<DllImport(Lib)>]
    extern int ps_batch_set_state(nativeint handle, int w0, int w1, float[] pos, float[] vel, float[] rot, float[] angMom)

    [<DllImport(Lib)>]
    extern int ps_batch_get_contacts(nativeint handle, int world, int& nPairs, int[] pairIds, int[] firstContact, float[] normal, float[] position, float[] penetration, uint32[] featureId, int pairCapacity, int contactCapacity)

    [<DllImport(Lib)>]
    extern int ps_batch_apply_impulse(nativeint handle, int world, int body, float[] impulse, float[] offset)

    [<DllImport(Lib)>]
    extern int ps_batch_add_body(nativeint handle, int world, PsBodiesView& one)

    [<DllImport(Lib)>]
    extern int ps_batch_reset(nativeint handle)

    /// error classes -> the exception types the reference throws on the same conditions
    let check rc =
        if rc <> 0 then
            let msg = Marshal.PtrToStringAnsi(ps_last_error ())

            match rc with
            | -1 -> invalidArg "native" msg // ArgumentException (RigidBodyPrototype.fs:61-67, SpatialTree.fs:171-172)
            | -2 -> raise (KeyNotFoundException msg) // bad world / body id (SimulatorState.fs:64,113-129)
            | -5 -> invalidOp msg // InvalidOperationException (Simulator.fs:124-125)
            | _ -> failwithf "ps_b200 error %d: %s" rc msg // CUDA failure / NaN / capacity: never a CPU reroute
