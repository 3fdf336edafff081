// NativeBatch.fs — marshalling between SimulatorState and the C ABI, and the additive BatchSimulator type.
// Goes between Native.fs and Simulator.fs in PhysicsSimulatorCore.fsproj (it needs SimulatorState's representation,
// which is private to the namespace, and nothing from Simulator.fs).  NOT COMPILED here: see Native.fs.
namespace PhysicsSimulator

open System
open System.Runtime.InteropServices
open PhysicsSimulator.Collisions
open PhysicsSimulator.Entities
open PhysicsSimulator.Utilities

/// SoA image of one or more SimulatorStates in the layout of ps_bodies_view: body index fastest inside a world,
/// worlds concatenated; ids are the list index (SimulatorState.fs:41-42), so body i of world w is Objects.[i].
type internal BodyArrays =
    { WorldOffset: int[]
      Shape: int[]
      Dims: float[]
      Mass: float[]
      InertiaInv: float[]
      Pos: float[]
      Vel: float[]
      Rot: float[]
      AngMom: float[]
      Force: float[]
      Torque: float[]
      Elasticity: float[]
      MuKinetic: float[]
      MuStatic: float[] }

module internal Marshalling =
    let private put3 (a: float[]) i (v: Vector3D) =
        a.[3 * i] <- v.X
        a.[3 * i + 1] <- v.Y
        a.[3 * i + 2] <- v.Z

    /// MathNet's DenseMatrix is column-major: the ABI wants m.[r, c] at rot.[9 i + 3 r + c]
    let private putRot (a: float[]) i (m: Matrix3) =
        let m = m.Get

        for r in 0..2 do
            for c in 0..2 do
                a.[9 * i + 3 * r + c] <- m.[r, c]

    let ofStates (states: SimulatorState[]) =
        let counts = states |> Array.map (fun s -> s.Objects.Count)
        let off = Array.scan (+) 0 counts
        let n = off.[off.Length - 1]

        let a =
            { WorldOffset = off
              Shape = Array.zeroCreate n
              Dims = Array.zeroCreate (3 * n)
              Mass = Array.zeroCreate n
              InertiaInv = Array.zeroCreate (3 * n)
              Pos = Array.zeroCreate (3 * n)
              Vel = Array.zeroCreate (3 * n)
              Rot = Array.zeroCreate (9 * n)
              AngMom = Array.zeroCreate (3 * n)
              Force = Array.zeroCreate (3 * n)
              Torque = Array.zeroCreate (3 * n)
              Elasticity = Array.zeroCreate n
              MuKinetic = Array.zeroCreate n
              MuStatic = Array.zeroCreate n }

        states
        |> Array.iteri (fun w state ->
            for KeyValue(id, obj) in state.Objects do
                let i = off.[w] + id.Get

                match obj.Collider with
                | Sphere s ->
                    a.Shape.[i] <- 0
                    a.Dims.[3 * i] <- s.Radius
                | Box b ->
                    a.Shape.[i] <- 1
                    a.Dims.[3 * i] <- b.XSize
                    a.Dims.[3 * i + 1] <- b.YSize
                    a.Dims.[3 * i + 2] <- b.ZSize

                match obj.PhysicalObject with
                | RigidBody rb ->
                    a.Mass.[i] <- rb.MassCenter.Mass.GetValue // Mass.Infinite -> +infinity (Entities/Base.fs:13-20)
                    let iinv = rb.PrincipalRotationalInertiaInverse.Get // the API only builds diagonal tensors (Box.fs:18-19, Sphere.fs:12)

                    for k in 0..2 do
                        a.InertiaInv.[3 * i + k] <- iinv.[k, k]

                    put3 a.Pos i rb.MassCenterPosition
                    put3 a.Vel i rb.MassCenterVelocity
                    putRot a.Rot i rb.Variables.Orientation
                    put3 a.AngMom i rb.Variables.AngularMomentum
                    a.Elasticity.[i] <- rb.ElasticityCoeff
                    a.MuKinetic.[i] <- rb.KineticFrictionCoeff
                    a.MuStatic.[i] <- rb.StaticFrictionCoeff
                | Particle _ -> invalidArg "state" "the native stepper takes rigid bodies (RigidBodyPrototype.build only makes those)"

                // gravity is already folded into ExternalForces by SimulatorStateBuilder (SimulatorState.fs:26-36,67-83)
                put3 a.Force i state.ExternalForces.[id]
                put3 a.Torque i state.ExternalTorques.[id])

        a

    let toConfig (step: StepConfiguration) (dt: TimeSpan) (broadPhase: BroadPhaseCollisionDetectionKind) recordContacts =
        let mutable c = PsStepConfig()
        c.epsilon <- step.Epsilon
        c.baumgarte <- step.BaumgarteTerm
        c.allowedPenetration <- step.AllowedPenetration
        c.maxCorrectionVelocity <- step.MaxCorrectionVelocity
        c.dt <- dt.TotalSeconds
        c.solverIterations <- step.CollisionSolverIterationCount
        c.enableFriction <- (if step.EnableFriction then 1 else 0)
        c.integratorKind <- (match step.RigidBodyIntegratorKind with | FirstOrder -> 0 | AugmentedSecondOrder -> 1)
        c.recordContacts <- (if recordContacts then 1 else 0)

        match broadPhase with
        | Dummy -> c.broadphaseKind <- 0
        | SpatialTree t ->
            c.broadphaseKind <- 1
            c.treeLeafCapacity <- t.LeafCapacity
            c.treeMaxDepth <- t.MaxDepth
            c.treeMinX <- t.SpaceBoundaries.Min.X
            c.treeMinY <- t.SpaceBoundaries.Min.Y
            c.treeMinZ <- t.SpaceBoundaries.Min.Z
            c.treeMaxX <- t.SpaceBoundaries.Max.X
            c.treeMaxY <- t.SpaceBoundaries.Max.Y
            c.treeMaxZ <- t.SpaceBoundaries.Max.Z

        c

    /// write the dynamic state the native step produced back into a SimulatorState (unmarshal)
    let withDynamicState (off: int) (pos: float[]) (vel: float[]) (rot: float[]) (angMom: float[]) (state: SimulatorState) =
        let v3 (a: float[]) i = Vector3D.create a.[3 * i] a.[3 * i + 1] a.[3 * i + 2]

        let objects =
            state.Objects
            |> Map.map (fun id obj ->
                let i = off + id.Get

                match obj.PhysicalObject with
                | RigidBody rb ->
                    let m = MathNet.Numerics.LinearAlgebra.Matrix<float>.Build.Dense(3, 3, fun r c -> rot.[9 * i + 3 * r + c])

                    { obj with
                        PhysicalObject =
                            RigidBody
                                { rb with
                                    MassCenter = { rb.MassCenter with Variables = { Position = v3 pos i; Velocity = v3 vel i } }
                                    Variables = { Orientation = Matrix3.ofMatrix m; AngularMomentum = v3 angMom i } } }
                | Particle _ -> obj)

        { state with Objects = objects }

/// Many independent worlds stepped together on one GPU (additive type; nothing existing is renamed).
/// Its executable specification is the Python class of the same name (physicssimulator_b200/simulator.py).
type BatchSimulator(worlds: RigidBodyPrototype list[], ?configuration0: Configuration, ?device: int) =
    let configuration = defaultArg configuration0 Configuration.getDefault

    let build () =
        worlds
        |> Array.map (fun protos ->
            protos
            |> SimulatorStateBuilder.fromPrototypes
            |> SimulatorStateBuilder.withConfiguration configuration.StepConfig
            |> SimulatorStateBuilder.withBroadPhase configuration.BroadPhaseCollisionDetectionKind)

    let mutable states = build ()
    let arrays = Marshalling.ofStates states
    let n = arrays.WorldOffset.[arrays.WorldOffset.Length - 1]
    let pos, vel, rot, angMom = Array.copy arrays.Pos, Array.copy arrays.Vel, Array.copy arrays.Rot, Array.copy arrays.AngMom
    let mutable handle = 0n

    do
        Native.check (Native.ps_init (defaultArg device -1))
        let pins = ResizeArray<GCHandle>()

        let pin (a: Array) =
            let h = GCHandle.Alloc(a, GCHandleType.Pinned)
            pins.Add h
            h.AddrOfPinnedObject()

        try
            let mutable view = PsBodiesView()
            view.nWorlds <- worlds.Length
            view.worldOffset <- pin arrays.WorldOffset
            view.shape <- pin arrays.Shape
            view.dims <- pin arrays.Dims
            view.mass <- pin arrays.Mass
            view.inertiaInv <- pin arrays.InertiaInv
            view.pos <- pin arrays.Pos
            view.vel <- pin arrays.Vel
            view.rot <- pin arrays.Rot
            view.angMom <- pin arrays.AngMom
            view.force <- pin arrays.Force
            view.torque <- pin arrays.Torque
            view.elasticity <- pin arrays.Elasticity
            view.muKinetic <- pin arrays.MuKinetic
            view.muStatic <- pin arrays.MuStatic

            let mutable cfg =
                Marshalling.toConfig configuration.StepConfig configuration.SimulationStepInterval configuration.BroadPhaseCollisionDetectionKind true

            Native.check (Native.ps_batch_create (&handle, &view, 0n, &cfg))
        finally
            for h in pins do
                h.Free()

    member _.NumWorlds = worlds.Length

    /// n x Simulator.updateSimulation (Simulator.fs:47-62) on every world; the state stays on the device
    member _.Step(nSteps: int) = Native.check (Native.ps_batch_step (handle, nSteps))

    /// read the device state back and return the worlds as SimulatorStates (AllPhysicalObjects per world)
    member _.States =
        Native.check (Native.ps_batch_get_state (handle, 0, worlds.Length, pos, vel, rot, angMom))
        states <- states |> Array.mapi (fun w s -> s |> Marshalling.withDynamicState arrays.WorldOffset.[w] pos vel rot angMom)
        states

    /// marshal -> n steps -> unmarshal in one native call (copies pipelined against the step)
    member this.StepHost(nSteps: int) =
        Native.check (Native.ps_batch_step_host (handle, nSteps, pos, vel, rot, angMom, pos, vel, rot, angMom))
        states <- states |> Array.mapi (fun w s -> s |> Marshalling.withDynamicState arrays.WorldOffset.[w] pos vel rot angMom)
        states

    member _.ApplyImpulse (world: int) (id: SimulatorObjectIdentifier) (impulse: Vector3D) (offset: Vector3D) =
        Native.check (Native.ps_batch_apply_impulse (handle, world, id.Get, [| impulse.X; impulse.Y; impulse.Z |], [| offset.X; offset.Y; offset.Z |]))

    member _.Reset() =
        Native.check (Native.ps_batch_reset handle)
        states <- build ()

    interface IDisposable with
        member _.Dispose() =
            if handle <> 0n then
                Native.ps_batch_destroy handle |> ignore
                handle <- 0n
