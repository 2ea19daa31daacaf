!  ref_golden.f90 -- TEST INFRASTRUCTURE.  Driver (our code) around the UNMODIFIED reference module
!  /root/reference/src/splpak.F90: reads problems from stdin, prints the reference's own coefficients and
!  splde values with 17 significant digits.  Built by `make -C oracle ref` into oracle/_ref/ when a Fortran
!  compiler exists (none does in this image); tests/test_reference_binary.py then pins the C oracle, the numpy
!  model and the CUDA path against these outputs and regenerates tests/golden/reference_*.json.
!
!  stdin, per problem:  ndim ndata nq weighted xtrap / nodes(ndim) / xmin(ndim) / xmax(ndim) /
!                       ndata lines: x(ndim) y [w] / nq lines: x(ndim) nderiv(ndim)
!  stdout:              "ierror" i / "coef" ncol values / "eval" nq lines: value ierror
program ref_golden
    use splpak_module
    implicit none
    integer :: ndim, ndata, nq, weighted, ncol, nwrk, ierror, i, ios
    real(splpak_wp) :: xtrap, v
    integer, allocatable :: nodes(:), nderiv(:)
    real(splpak_wp), allocatable :: xmin(:), xmax(:), xdata(:,:), ydata(:), wdata(:), coef(:), work(:), x(:)
    type(splpak_type) :: s
    do
        read(*,*,iostat=ios) ndim, ndata, nq, weighted, xtrap
        if (ios /= 0) exit
        allocate(nodes(ndim), nderiv(ndim), xmin(ndim), xmax(ndim), x(ndim))
        read(*,*) nodes
        read(*,*) xmin
        read(*,*) xmax
        allocate(xdata(ndim,ndata), ydata(ndata), wdata(ndata))
        do i = 1, ndata
            if (weighted /= 0) then
                read(*,*) xdata(:,i), ydata(i), wdata(i)
            else
                read(*,*) xdata(:,i), ydata(i)
            end if
        end do
        ncol = product(nodes)
        nwrk = ncol*(ncol+1) + 1
        allocate(coef(ncol), work(nwrk))
        if (weighted /= 0) then
            call s%initialize(ndim,xdata,ndim,ydata,wdata,ndata,xmin,xmax,nodes,xtrap,coef,ncol,work,nwrk,ierror)
        else
            call s%initialize(ndim,xdata,ndim,ydata,ndata,xmin,xmax,nodes,xtrap,coef,ncol,work,nwrk,ierror)
        end if
        write(*,'(a,1x,i0)') 'ierror', ierror
        write(*,'(a,1x,i0)') 'coef', ncol
        do i = 1, ncol
            write(*,'(es25.17e3)') coef(i)
        end do
        write(*,'(a,1x,i0)') 'eval', nq
        do i = 1, nq
            read(*,*) x, nderiv
            v = s%evaluate(ndim,x,nderiv,coef,xmin,xmax,nodes,ierror)
            write(*,'(es25.17e3,1x,i0)') v, ierror
        end do
        deallocate(nodes, nderiv, xmin, xmax, x, xdata, ydata, wdata, coef, work)
    end do
end program ref_golden
