/*
 * WideFieldModelB200 -- drop-in sibling of microTiPi.epifluorescence.WideFieldModel whose arithmetic runs in
 * libwfm_b200.so (include/wfm_b200.h) on a B200.  SURVEY.md section 8, rows b1 / f2.
 *
 * STATUS: NOT COMPILED IN THE BUILD CONTAINER (no JDK, no TiPi / JTransforms jars there).  It is written against
 *   - the C ABI of include/wfm_b200.h (every downcall below names its entry point), and
 *   - the public surface of the reference class, member by member (WFM:<line> = WideFieldModel.java, MM:<line> =
 *     MicroscopeModel.java), with the same argument meaning and the same IllegalArgumentException sites.
 * The same ABI calls, in the same order, are what microtipi_b200/wide_field_model.py (ctypes) and
 * include/wfm_b200.hpp (C++) make, and those two ARE exercised by the test-suite.
 *
 * Requirements on the JVM side: JDK >= 22 (java.lang.foreign), TiPi on the class path, libwfm_b200.so on
 * java.library.path / LD_LIBRARY_PATH, run with --enable-native-access=ALL-UNNAMED.
 *
 * Placement: same package as WideFieldModel, extends the UNCHANGED MicroscopeModel: PSF_Estimation reads the protected
 * parameterCoefs[] and calls the protected computePsf() through a MicroscopeModel-typed reference
 * (PSF_Estimation.java:117,204), which is legal only inside this package hierarchy.
 */
package microTiPi.epifluorescence;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

import microTiPi.microscopy.MicroscopeModel;
import mitiv.array.Array3D;
import mitiv.array.Array4D;
import mitiv.array.Double3D;
import mitiv.array.Double4D;
import mitiv.array.Float3D;
import mitiv.array.Float4D;
import mitiv.base.Shape;
import mitiv.linalg.shaped.DoubleShapedVector;
import mitiv.linalg.shaped.DoubleShapedVectorSpace;
import mitiv.linalg.shaped.FloatShapedVector;
import mitiv.linalg.shaped.ShapedVector;
import mitiv.linalg.shaped.ShapedVectorSpace;

public class WideFieldModelB200 extends MicroscopeModel implements AutoCloseable {

    // ---- the C ABI ---------------------------------------------------------------------------------------------
    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.mapLibraryName("wfm_b200"), Arena.global());

    private static MethodHandle fn(String name, FunctionDescriptor d) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), d);
    }
    private static FunctionDescriptor ints(java.lang.foreign.MemoryLayout... args) { return FunctionDescriptor.of(JAVA_INT, args); }
    /**
     * Downcall that may be handed HEAP segments (MemorySegment.ofArray): Linker.Option.critical(true) pins the Java array
     * for the duration of the call and passes its address, so psf / q / cpxPsf move between the TiPi arrays (double[] /
     * float[] on the Java heap, SURVEY 8 b4) and the device without a Java-side copy.  The library sees ordinary pageable
     * memory and stages it itself (host threads + pinned slots, wfm_api.cu "pageable host arrays"): measured 27 ms per
     * 2 x 537 MB step against 78 ms for a plain cudaMemcpy of pageable memory and 11-12 ms from pinned buffers.
     * The callee makes no upcalls and returns within milliseconds, as a critical call must.
     */
    private static MethodHandle fnHeap(String name, FunctionDescriptor d) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), d, Linker.Option.critical(true));
    }

    private static final MethodHandle CREATE = fn("wfm_create", ints(ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_INT, JAVA_INT));
    private static final MethodHandle CREATE_MULTI = fn("wfm_create_multi", ints(ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_INT, ADDRESS, JAVA_INT));
    private static final MethodHandle DESTROY = fn("wfm_destroy", ints(ADDRESS));
    private static final MethodHandle LAST_ERROR = fn("wfm_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    private static final MethodHandle SET_OPTICS = fn("wfm_set_optics", ints(ADDRESS, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_DOUBLE));
    private static final MethodHandle BUILD_BASIS = fn("wfm_build_basis", ints(ADDRESS, JAVA_INT, JAVA_INT));
    private static final MethodHandle GET_BASIS = fn("wfm_get_basis", ints(ADDRESS, ADDRESS, JAVA_INT));
    private static final MethodHandle SET_PHASE = fn("wfm_set_phase", ints(ADDRESS, ADDRESS, JAVA_INT));
    private static final MethodHandle SET_MODULUS = fn("wfm_set_modulus", ints(ADDRESS, ADDRESS, JAVA_INT));
    private static final MethodHandle SET_DEFOCUS = fn("wfm_set_defocus", ints(ADDRESS, ADDRESS, JAVA_INT));
    private static final MethodHandle GET_RHO = fn("wfm_get_rho", ints(ADDRESS, ADDRESS));
    private static final MethodHandle GET_PHI = fn("wfm_get_phi", ints(ADDRESS, ADDRESS));
    private static final MethodHandle GET_PSI = fn("wfm_get_psi", ints(ADDRESS, ADDRESS));
    private static final MethodHandle GET_MASK = fn("wfm_get_mask", ints(ADDRESS, ADDRESS));
    private static final MethodHandle COMPUTE_PSF = fn("wfm_compute_psf", ints(ADDRESS));
    private static final MethodHandle INVALIDATE = fn("wfm_invalidate", ints(ADDRESS));
    private static final MethodHandle GET_PSF = fnHeap("wfm_get_psf", ints(ADDRESS, ADDRESS));
    private static final MethodHandle GET_CPX = fnHeap("wfm_get_cpx_psf", ints(ADDRESS, ADDRESS));
    private static final MethodHandle GET_MTF = fnHeap("wfm_get_mtf", ints(ADDRESS, ADDRESS));
    private static final MethodHandle APPLY_J = fnHeap("wfm_apply_jacobian", ints(ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
    private static final MethodHandle HOST_ALLOC = fn("wfm_host_alloc", ints(ADDRESS, java.lang.foreign.ValueLayout.JAVA_LONG));
    private static final MethodHandle HOST_FREE = fn("wfm_host_free", ints(ADDRESS));

    private static final int WFM_ERR_INVALID_ARG = -1;

    // ---- reference constants (WFM:113-123) ------------------------------------------------------------------------
    public static final int DEFOCUS = 0;
    public static final int PHASE = 1;
    public static final int MODULUS = 2;
    public static final int[] parametersFlag = {DEFOCUS, PHASE, MODULUS};

    // ---- state mirrored on the Java side (WFM:81-108); the arrays themselves live on the device ---------------------
    protected double lambda, ni, NA, lambda_ni, radius;
    protected double deltaX = 0.0, deltaY = 0.0;
    protected int nModulus, nPhase, Nzern;
    protected boolean radial;

    private final Arena arena = Arena.ofShared();
    private MemorySegment handle = MemorySegment.NULL;
    private final long vox, elemBytes;
    // (no staging buffers on the Java side: psf / q / cpxPsf are handed over as heap segments, see fnHeap)

    /** WFM:137-152 (no Zernike modes: nPhase = 0, nModulus = 1). */
    public WideFieldModelB200(Shape psfShape, double NA, double lambda, double ni, double dxy, double dz, boolean radial, boolean single) {
        this(psfShape, 0, 1, NA, lambda, ni, dxy, dz, radial, single);
    }

    /** WFM:154-188. */
    public WideFieldModelB200(Shape psfShape, int nPhase, int nModulus, double NA, double lambda, double ni,
                              double dxy, double dz, boolean radial, boolean single) {
        super(psfShape, dxy, dz, single);                                           // MM:62-78
        if (Nx != Ny) throw new IllegalArgumentException("Nx should equal Ny");     // WFM:158-160
        this.lambda = lambda; this.ni = ni; this.NA = NA; this.radial = radial;
        this.Nzern = 4;                                                             // WFM:163
        this.radius = NA / lambda;                                                  // WFM:165
        this.lambda_ni = ni / lambda;                                               // WFM:166
        MemorySegment out = arena.allocate(ADDRESS);
        int[] devs = deviceList();
        if (devs.length > 1)        // the stack over several GPUs behind the SAME calls: one z-slab per device (wfm_create_multi)
            check(MemorySegment.NULL, call(CREATE_MULTI, out, Nx, Ny, Nz, dxy, dz, single ? 1 : 0,
                                           arena.allocateFrom(JAVA_INT, devs), devs.length));
        else
            check(MemorySegment.NULL, call(CREATE, out, Nx, Ny, Nz, dxy, dz, single ? 1 : 0, devs[0]));
        handle = out.get(ADDRESS, 0);
        vox = (long) Nx * Ny * Nz;
        elemBytes = single ? 4 : 8;
        check(handle, call(SET_OPTICS, handle, NA, lambda, ni));                    // computeMaskPupil()  WFM:174, 1374-1406
        this.nModulus = Math.max(1, nModulus);                                      // WFM:176-179
        this.nPhase = nPhase;
        parameterSpace = new DoubleShapedVectorSpace[3];                            // WFM:181-182
        parameterCoefs = new DoubleShapedVector[3];
        setNModulus();                                                              // WFM:185
        setNPhase();                                                                // WFM:186
        setDefocus(new double[] {ni / lambda, deltaX, deltaY});                     // WFM:187, 1562-1564
    }

    /** Devices this model lives on: -Dwfm.devices=0,1,...,7 spreads the z-planes over the GPUs of the box (the caller
     *  stays one thread, PSF_Estimation.java:202-217: every call below fans out inside the library); default -Dwfm.device=0. */
    protected int[] deviceList() {
        String list = System.getProperty("wfm.devices");
        if (list == null || list.isBlank()) return new int[] {Integer.getInteger("wfm.device", 0)};
        return java.util.Arrays.stream(list.split(",")).map(String::trim).mapToInt(Integer::parseInt).toArray();
    }

    // ---- the hot path ------------------------------------------------------------------------------------------------
    /** WFM:206-396.  One pipeline launch for all z-planes; no-op when the PSF is valid (WFM:207). */
    @Override
    public void computePsf() {
        if (PState > 0) return;
        check(handle, call(COMPUTE_PSF, handle));
        PState = 1;                                                                 // WFM:395
    }

    /** WFM:399-409: dispatch on the IDENTITY of the vector space. */
    @Override
    public DoubleShapedVector apply_Jacobian(ShapedVector grad, ShapedVectorSpace xspace) {
        if (xspace == parameterSpace[DEFOCUS]) return apply_J_defocus(grad);
        if (xspace == parameterSpace[PHASE]) return apply_J_phase(grad);
        if (xspace == parameterSpace[MODULUS]) return apply_J_modulus(grad);
        throw new IllegalArgumentException("DoubleShapedVector grad does not belong to any space");
    }

    /** WFM:429-730 (quirk Q1: the intended sum over z; wfm_set_modulus_mode selects the live last-plane behaviour). */
    public DoubleShapedVector apply_J_modulus(final ShapedVector q) { return applyJ(MODULUS, q); }
    /** WFM:738-1021. */
    public DoubleShapedVector apply_J_phase(ShapedVector q) { return applyJ(PHASE, q); }
    /** WFM:1029-1369 (the live half-gradient, quirk Q3). */
    public DoubleShapedVector apply_J_defocus(ShapedVector q) { return applyJ(DEFOCUS, q); }

    private DoubleShapedVector applyJ(int flag, ShapedVector q) {
        DoubleShapedVectorSpace space = parameterSpace[flag];
        if (space == null) throw new IllegalArgumentException("DoubleShapedVector grad does not belong to any space");
        int n = space.getNumber();
        // q straight from the TiPi vector's backing array (no copy on the Java side)
        MemorySegment qs = isSingle() ? MemorySegment.ofArray(((FloatShapedVector) q).getData())
                                      : MemorySegment.ofArray(((DoubleShapedVector) q).getData());
        if (qs.byteSize() != vox * elemBytes) throw new IllegalArgumentException("q does not have the shape of the PSF");
        double[] out = new double[n];
        check(handle, call(APPLY_J, handle, flag, qs, MemorySegment.ofArray(out), n));   // recomputes the PSF if dirty (quirk Q5)
        PState = 1;
        return space.wrap(out);
    }

    /** WFM:412-422. */
    @Override
    public void setParam(DoubleShapedVector param) {
        if (param.getOwner() == parameterSpace[DEFOCUS]) setDefocus(param);
        else if (param.getOwner() == parameterSpace[PHASE]) setPhase(param);
        else if (param.getOwner() == parameterSpace[MODULUS]) setModulus(param);
        else throw new IllegalArgumentException("DoubleShapedVector param does not belong to any space");
    }
    /** WFM:1553-1556. */
    @Override
    public void setParam(double[] param) { setDefocus(param); }

    // ---- pupil setters ------------------------------------------------------------------------------------------------
    /** WFM:1452-1499: psi and maskPupil from {ni/lambda, deltaX, deltaY}. */
    public void computeDefocus() { nativeVector(SET_DEFOCUS, new double[] {lambda_ni, deltaX, deltaY}); }

    /** WFM:1510-1534. */
    public void setDefocus(DoubleShapedVector defoc) {
        if (!defoc.belongsTo(parameterSpace[DEFOCUS]))
            throw new IllegalArgumentException("defocus  does not belong to the parameterSpace[DEFOCUS]");
        int n = defoc.getNumber();
        if (n != 1 && n != 3) throw new IllegalArgumentException("bad defocus  parameters");   // WFM:1530 (+ quirk Q4: n == 2)
        parameterCoefs[DEFOCUS] = defoc;
        if (n == 3) { deltaX = defoc.get(1); deltaY = defoc.get(2); }
        lambda_ni = defoc.get(0);
        ni = lambda_ni * lambda;                                                    // WFM:1523
        nativeVector(SET_DEFOCUS, defoc.getData());
        freeMem();
    }
    /** WFM:1543-1549. */
    public void setDefocus(double[] defoc) {
        if (parameterSpace[DEFOCUS] == null) parameterSpace[DEFOCUS] = new DoubleShapedVectorSpace(3);
        setDefocus(parameterSpace[DEFOCUS].wrap(defoc));
    }
    /** WFM:1573-1579. */
    public void setPupilAxis(double[] axis) { setDefocus(new double[] {ni / lambda, axis[0], axis[1]}); }
    /** WFM:1698-1707. */
    public void setNi(Double value) { ni = value; lambda_ni = ni / lambda; setDefocus(new double[] {ni / lambda, deltaX, deltaY}); }

    /** WFM:1588-1610. */
    public void setModulus(DoubleShapedVector modulus) {
        if (!modulus.belongsTo(parameterSpace[MODULUS]))
            throw new IllegalArgumentException("DoubleShapedVector beta does not belong to the modulus space");
        parameterCoefs[MODULUS] = modulus;
        nativeVector(SET_MODULUS, modulus.getData());
        freeMem();
    }
    /** WFM:1616-1620. */
    public void setModulus(double[] modulus) { setNModulus(modulus.length); setModulus(parameterSpace[MODULUS].wrap(modulus)); }

    /** WFM:1625-1649. */
    public void setPhase(DoubleShapedVector phase) {
        if (parameterSpace[PHASE] == null || !phase.belongsTo(parameterSpace[PHASE]))
            throw new IllegalArgumentException("phase parameter does not belong to the right space  ");
        parameterCoefs[PHASE] = phase;
        nativeVector(SET_PHASE, phase.getData());
        freeMem();
    }
    /** WFM:1655-1665. */
    public void setPhase(double[] alpha) {
        if (alpha == null || alpha.length == 0) {
            nPhase = 0; parameterSpace[PHASE] = null; parameterCoefs[PHASE] = null;
            check(handle, call(SET_PHASE, handle, MemorySegment.NULL, 0));          // the device side drops its phase vector too
            freeMem();
            return;
        }
        setNPhase(alpha.length);
        setPhase(parameterSpace[PHASE].wrap(alpha));
    }

    /** WFM:1919-1922 / 1899-1914. */
    public void setNPhase(int nPh) { nPhase = nPh; setNPhase(); }
    private void setNPhase() {
        if (nPhase > 0) {
            parameterSpace[PHASE] = new DoubleShapedVectorSpace(nPhase);
            Nzern = Math.max(nPhase + (radial ? 1 : 3), parameterSpace[MODULUS].getNumber());
            computeZernike();
            parameterCoefs[PHASE] = parameterSpace[PHASE].create(0.);
            setPhase(parameterCoefs[PHASE]);
        } else {
            parameterSpace[PHASE] = null;
            parameterCoefs[PHASE] = null;
        }
    }
    /** WFM:1930-1934 / 1939-1961. */
    public void setNModulus(int nMod) { nModulus = nMod; setNModulus(); }
    private void setNModulus() {
        if (nModulus < 1) nModulus = 1;
        parameterSpace[MODULUS] = new DoubleShapedVectorSpace(nModulus);
        Nzern = parameterSpace[PHASE] == null ? nModulus
                                              : Math.max(parameterSpace[PHASE].getNumber() + (radial ? 1 : 3), nModulus);
        computeZernike();
        parameterCoefs[MODULUS] = parameterSpace[MODULUS].create(0.);
        parameterCoefs[MODULUS].set(0, 1.);
        setModulus(parameterCoefs[MODULUS]);
    }
    /** WFM:194-197: Zernike.zernikeArray + Gram-Schmidt, on the device. */
    private void computeZernike() { check(handle, call(BUILD_BASIS, handle, Nzern, radial ? 1 : 0)); }

    // ---- getters (each recomputes the PSF when dirty, like WFM:1674-1676 etc.) -------------------------------------------
    public double[] getRho() { return pupil(GET_RHO); }                             // WFM:1673
    public double[] getPhi() { return pupil(GET_PHI); }                             // WFM:1713
    public double[] getPsi() { return pupil(GET_PSI); }                             // WFM:1723
    public boolean[] getMaskPupil() {                                               // WFM:1784
        if (PState < 1) computePsf();
        try (Arena a = Arena.ofConfined()) {
            MemorySegment m = a.allocate((long) Nx * Ny);
            check(handle, call(GET_MASK, handle, m));
            boolean[] out = new boolean[Nx * Ny];
            for (int i = 0; i < out.length; ++i) out[i] = m.get(JAVA_BYTE, i) != 0;
            return out;
        }
    }
    public double getLambda() { return lambda; }                                    // WFM:1684
    public double getNi() { return ni; }                                            // WFM:1691
    public DoubleShapedVector getModulusCoefs() { return parameterCoefs[MODULUS]; } // WFM:1736
    public DoubleShapedVector getPhaseCoefs() { return parameterCoefs[PHASE]; }     // WFM:1743
    public double[] getDefocusMultiplyByLambda() {                                  // WFM:1750-1756
        if (PState < 1) computePsf();
        return new double[] {lambda_ni * lambda, deltaX * lambda, deltaY * lambda};
    }
    public double[] getDefocus() {                                                  // WFM:1761-1767
        if (PState < 1) computePsf();
        return new double[] {lambda_ni, deltaX, deltaY};
    }
    public double[] getPupilShift() {                                               // WFM:1772-1778
        if (PState < 1) computePsf();
        return new double[] {deltaX, deltaY};
    }
    public int getNZern() { return Nzern; }                                         // WFM:1841
    public int getNModulus() { return parameterCoefs[MODULUS].getNumber(); }        // WFM:1981
    public int getNPhase() { return parameterCoefs[PHASE] == null ? 0 : parameterCoefs[PHASE].getNumber(); }   // WFM:1988
    @Override public int[] getParametersFlags() { return parametersFlag; }          // WFM:2000

    /** WFM:1834 / 1849: the orthonormal basis, Nzern planes of Nx*Ny. */
    public double[] getZernike() {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment z = a.allocate(JAVA_DOUBLE, (long) Nzern * Nx * Ny);
            check(handle, call(GET_BASIS, handle, z, Nzern));
            return z.toArray(JAVA_DOUBLE);
        }
    }
    public double[] getZernike(int k) { return java.util.Arrays.copyOfRange(getZernike(), k * Nx * Ny, (k + 1) * Nx * Ny); }

    /** WFM:1798-1804. */
    @Override
    public Array3D getPsf() {
        if (PState < 1) computePsf();
        if (isSingle()) {
            float[] a = new float[(int) vox];
            check(handle, call(GET_PSF, handle, MemorySegment.ofArray(a)));          // device -> the array TiPi will own
            psf = Float3D.wrap(a, psfShape);
        } else {
            double[] a = new double[(int) vox];
            check(handle, call(GET_PSF, handle, MemorySegment.ofArray(a)));
            psf = Double3D.wrap(a, psfShape);
        }
        return psf;
    }
    /** WFM:1856-1861: conj(FFT2(A_z)), shape (2, Nx, Ny, Nz). */
    public Array4D get_cpxPsf() {
        if (PState < 1) computePsf();
        Shape s = new Shape(2, Nx, Ny, Nz);
        if (isSingle()) {
            float[] a = new float[(int) (2 * vox)];
            check(handle, call(GET_CPX, handle, MemorySegment.ofArray(a)));
            return Float4D.wrap(a, s);
        }
        double[] a = new double[(int) (2 * vox)];
        check(handle, call(GET_CPX, handle, MemorySegment.ofArray(a)));
        return Double4D.wrap(a, s);
    }
    /** WFM:1807-1828 as intended (the reference copy loop `i = i++` never terminates, quirk Q8): FFT3 of the PSF. */
    @Override
    public Array4D getMtf() {
        double[] a = new double[(int) (2 * vox)];
        check(handle, call(GET_MTF, handle, MemorySegment.ofArray(a)));
        PState = 1;
        return Double4D.wrap(a, new Shape(2, Nx, Ny, Nz));
    }
    /** WFM:1866-1895 prints statistics of the pupil; kept as a one-line summary. */
    public void getInfo() {
        System.out.println("WideFieldModelB200 " + Nx + "x" + Ny + "x" + Nz + " NA=" + NA + " lambda=" + lambda + " ni=" + ni
                + " nPhase=" + getNPhase() + " nModulus=" + getNModulus() + " Nzern=" + Nzern);
    }

    /** WFM:1970-1974: PState = 0 (device buffers are kept for reuse). */
    @Override
    public void freeMem() {
        PState = 0;
        psf = null;
        if (!handle.equals(MemorySegment.NULL)) call(INVALIDATE, handle);
    }

    @Override
    public void close() {
        if (!handle.equals(MemorySegment.NULL)) { call(DESTROY, handle); handle = MemorySegment.NULL; }
        arena.close();
    }

    // ---- plumbing -------------------------------------------------------------------------------------------------------
    private double[] pupil(MethodHandle getter) {
        if (PState < 1) computePsf();
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(JAVA_DOUBLE, (long) Nx * Ny);
            check(handle, call(getter, handle, out));
            return out.toArray(JAVA_DOUBLE);
        }
    }
    private void nativeVector(MethodHandle setter, double[] v) {
        try (Arena a = Arena.ofConfined()) {
            check(handle, call(setter, handle, a.allocateFrom(JAVA_DOUBLE, v), v.length));
        }
    }
    private MemorySegment pinned(long bytes) {
        MemorySegment out = arena.allocate(ADDRESS);
        check(handle, call(HOST_ALLOC, out, bytes));
        return out.get(ADDRESS, 0).reinterpret(bytes);
    }
    private static int call(MethodHandle m, Object... args) {
        try { return (int) m.invokeWithArguments(args); }
        catch (Throwable t) { throw new IllegalStateException(t); }
    }
    /** Status -> exception: WFM_ERR_INVALID_ARG is the reference's IllegalArgumentException; nothing is swallowed (Q7). */
    private static void check(MemorySegment h, int rc) {
        if (rc == 0) return;
        String msg;
        try { msg = ((MemorySegment) LAST_ERROR.invokeWithArguments(h)).reinterpret(512).getString(0); }
        catch (Throwable t) { msg = "status " + rc; }
        if (rc == WFM_ERR_INVALID_ARG) throw new IllegalArgumentException(msg);
        throw new IllegalStateException("wfm_b200: " + msg + " (status " + rc + ")");
    }
}
